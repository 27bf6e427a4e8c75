#!/usr/bin/env python3
"""Per-kernel SASS inventory of avif-format_b200/lib/libavifgpu.so: instruction count and the mnemonics that show how
a kernel moves data (UBLKCP = cp.async.bulk copy-engine fetch, SYNCS = mbarrier, VIADDMNMX = DPX add-clamp, ...).

    python profiles/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "avif-format_b200", "lib", "libavifgpu.so")
WATCH = ["UBLKCP", "SYNCS", "LDG", "STG", "LDS", "STS", "VIADDMNMX", "VIMNMX", "DFMA", "DMUL", "DADD", "F2F", "MUFU", "FCHK", "FFMA2", "FMUL2", "FADD2", "R2P", "SHFL", "BAR", "BSSY", "CALL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    counts, ops, order, current = collections.Counter(), collections.defaultdict(collections.Counter), [], None
    index = 0
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            current = names[index]
            index += 1
            current = re.sub(r"\(anonymous namespace\)::", "", current)
            current = re.sub(r"\(.*", "", current).replace("void avifgpu::", "")
            order.append(current)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and current:
            counts[current] += 1
            ops[current][m.group(1)] += 1
    print("# cuobjdump -sass of lib/libavifgpu.so (sm_100a): instructions per kernel, and the data-movement mnemonics")
    for name in sorted(order):
        marks = " ".join(f"{op}={ops[name][op]}" for op in WATCH if ops[name][op])
        print(f"{name:70s} {counts[name]:6d}  {marks}")


if __name__ == "__main__":
    main()
