"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU, exports exactly what
include/avifgpu.h declares, refuses loudly to run without a device, and its host arithmetic (coefficients,
tables, geometry, validation) agrees with the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import cases
from avifgpu import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "avifgpu.h")).read()
    return sorted(set(re.findall(r"AVIFGPU_EXPORT\s+[\w\s\*]+?\b(avifgpu_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    import avifgpu
    assert declared_symbols() == sorted(avifgpu.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    import avifgpu
    lib = avifgpu.library()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.avifgpu_api_version() == abi.API_VERSION
    out = subprocess.run(["nm", "-D", "--defined-only", avifgpu.LIBRARY_PATH], capture_output=True, text=True).stdout
    exported = sorted(line.split()[-1] for line in out.splitlines() if " T " in line)
    assert exported == declared_symbols(), "the library must export the C ABI and nothing else"


def test_struct_sizes_match_the_c_compiler(tmp_path):
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "avifgpu.h"\nint main(void){printf("%zu %zu %zu %zu\\n",'
                   'sizeof(avifgpu_nclx),sizeof(avifgpu_planes),sizeof(avifgpu_encode_desc),sizeof(avifgpu_decode_desc));return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(abi.Nclx), C.sizeof(abi.Planes), C.sizeof(abi.EncodeDesc), C.sizeof(abi.DecodeDesc)]


def test_no_cpu_fallback_without_a_device():
    import torch
    import avifgpu
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(avifgpu.AvifGpuError) as info:
        avifgpu.Context(0)
    assert info.value.status == abi.ERR_NO_DEVICE
    assert "no CPU fallback" in info.value.message


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "avif-format_b200")
    for folder, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".cu", ".cuh", ".cpp", ".h", ".py", ".inc")) or name == "Makefile":
                text = open(os.path.join(folder, name), errors="replace").read()
                assert "liboracle" not in text and "libavifref" not in text and "avif_oracle" not in text, os.path.join(folder, name)


def test_host_arithmetic_matches_oracle(port):
    import avifgpu
    for matrix in (0, 1, 2, 4, 5, 6, 7, 9, 12):
        for primaries in (1, 4, 5, 6, 7, 8, 9, 10, 11, 12, 22, 2):
            n = abi.Nclx(1, primaries, 13, matrix, 1)
            assert cases.same_bits(avifgpu.yuv_coefficients(n), port.yuv_coefficients(n))
    assert cases.same_bits(avifgpu.yuv_coefficients(None), port.yuv_coefficients(None))
    for primaries in (1, 5, 6, 9):
        assert cases.same_bits(avifgpu.hlg_luma_coefficients(primaries), port.hlg_luma_coefficients(primaries))
    with pytest.raises(avifgpu.AvifGpuError):
        avifgpu.hlg_luma_coefficients(12)
    for depth in (8, 10, 12, 16):
        for full in (0, 1):
            for matrix in (0, 6, 9):
                for mono in (False, True):
                    n = abi.Nclx(1, 9, 16, matrix, full)
                    for a, b in zip(avifgpu.yuv_tables(n, depth, mono), port.yuv_tables(n, depth, mono)):
                        assert (a is None) == (b is None)
                        if a is not None:
                            assert cases.same_bits(a, b), (depth, full, matrix, mono)


def test_plane_geometry_matches_python_mirror():
    import avifgpu
    lib = avifgpu.library()
    w, h, b = C.c_int32(), C.c_int32(), C.c_int32()
    for _, desc, _, _ in cases.encode_cases([(37, 23), (1, 1)], full=True):
        shapes = abi.encode_plane_shapes(desc)
        assert lib.avifgpu_encode_host_col_bytes(C.byref(desc)) == desc.host_channels * (desc.host_depth // 8)
        for k in range(4):
            present = lib.avifgpu_encode_plane_geometry(C.byref(desc), k, C.byref(w), C.byref(h), C.byref(b))
            if shapes[k] is None:
                assert present == 0
            else:
                assert present == 1 and (h.value, w.value) == shapes[k] and b.value == (2 if desc.image_bit_depth > 8 else 1)
    for _, desc, _, _ in cases.decode_cases([(37, 23)], full=True):
        shapes = abi.decode_plane_shapes(desc)
        assert lib.avifgpu_decode_host_col_bytes(C.byref(desc)) == abi.decode_host_channels(desc) * (desc.host_depth // 8)
        for k in range(4):
            present = lib.avifgpu_decode_plane_geometry(C.byref(desc), k, C.byref(w), C.byref(h), C.byref(b))
            assert present == (0 if shapes[k] is None else 1)
            if shapes[k] is not None:
                assert (h.value, w.value) == shapes[k]
