"""What the committed float decode kernel's machine code looks like -- the claims DESIGN.md section 4.5 makes about it, checked on the
built library with cuobjdump (no GPU needed): per pixel 44 FP64 instructions for HLG + OOTF and 102 for PQ (glibc's sequences and
nothing more), and no branch-and-call division (FCHK ... CALL) inside the main loop of the PQ kernel with the verified quotient."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "avif-format_b200", "lib", "libavifgpu.so")


def main_loops():
    """{demangled kernel name: Counter of opcodes inside the kernel's largest backward-branch span}"""
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    loops = {}
    for block in sass.split("Function : ")[1:]:
        mangled = block.split("\n", 1)[0].strip()
        if "DecodeYccToRgbF32Kernel" not in mangled:
            continue
        instructions = [(int(m.group(1), 16), m.group(2).strip()) for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", block)]
        best = None
        for address, text in instructions:
            target = re.search(r"BRA\S*\s+.*?(0x[0-9a-f]+)", text)
            if target and int(target.group(1), 16) < address:
                span = address - int(target.group(1), 16)
                if best is None or span > best[0]:
                    best = (span, int(target.group(1), 16), address)
        counts = collections.Counter()
        for address, text in instructions:
            if best and best[1] <= address <= best[2]:
                words = text.split()
                if words[0].startswith("@"):
                    words = words[1:]
                counts[words[0].split(".")[0]] += 1
        loops[mangled] = counts
    return loops


@pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB), reason="needs cuobjdump and the built library")
def test_float_decode_loops_hold_glibcs_fp64_sequences_and_nothing_more():
    loops = main_loops()
    # template arguments <XS, YS, TRANSFER, ALPHA, FASTDIV>: 4:2:0 without alpha; TRANSFER 1 = HLG, 0 = PQ (include/avifgpu.h)
    hlg = next(v for k, v in loops.items() if "ILi1ELi1ELi1ELi0ELi0EE" in k)
    pq = next(v for k, v in loops.items() if "ILi1ELi1ELi0ELi0ELi1EE" in k)
    pq_ieee = next(v for k, v in loops.items() if "ILi1ELi1ELi0ELi0ELi0EE" in k)
    fp64 = lambda c: c["DFMA"] + c["DMUL"] + c["DADD"]  # noqa: E731
    pixels_per_iteration = 8  # a lane's 4 pixels of each row of a row pair
    assert fp64(hlg) == 44 * pixels_per_iteration    # 3 expf x 9 + 1 powf x 17 (two fewer than glibc's 19: the exponent-folded table)
    assert fp64(pq) == 102 * pixels_per_iteration    # 6 powf x 17
    assert fp64(pq_ieee) == 102 * pixels_per_iteration
    assert pq["MUFU"] >= 24 and pq_ieee["FCHK"] >= 24  # one reciprocal seed / one checked division per channel sample
    # the verified quotient leaves no checked division in the loop beyond the green-term fallback (2 chroma sites, taken only
    # when VerifyGreenDivision has not passed for the configuration)
    assert pq["FCHK"] <= 2 and pq["CALL"] <= 2
    assert hlg["FCHK"] <= 2 and hlg["CALL"] <= 2
    # packed FP32 and packed code clamps are what the source asks for
    assert hlg["FFMA2"] > 0 and hlg["FMUL2"] > 0 and hlg["VIMNMX"] >= 6
