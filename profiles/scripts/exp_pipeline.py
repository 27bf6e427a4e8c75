#!/usr/bin/env python3
"""Experiment: where the plug-in shuttle's time goes relative to the single host-pointer call (config 2, one B200).
Variants of the same 8K frame through the C ABI:
  A  one synchronous call, rows and planes in torch-pinned memory                       (= bench.py's e2e)
  A' the same with rows and planes from avifgpu_host_alloc (pinned on the GPU's NUMA node)
  B  six asynchronous calls of 728 rows (the shuttle's blocks), pinned rows and planes, one wait at the end
  C  as B with pageable planes (numpy): what the shuttle does with libheif's planes
  D  one synchronous call with pageable planes
Prints ms per image (mean of `steps`, and best)."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "avif-format_b200", "python"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import avifgpu  # noqa: E402
from avifgpu import abi  # noqa: E402

W, H = 7680, 4320
steps = 10
ctx = avifgpu.Context(0)
nclx = abi.Nclx(1, abi.PRIMARIES_BT2020, abi.TRANSFER_CHAR_PQ, abi.MATRIX_BT2020_NCL, 1)
desc = abi.EncodeDesc(W, H, 32, 3, abi.ALPHA_NONE, 12, abi.TRANSFER_PQ, 80, abi.LAYOUT_PLANAR_YCBCR, abi.CHROMA_420, abi.DOWN_FILTER_BOX, abi.GRAY16_LUT, nclx)
ctx.prepare_encode(desc)
shapes = abi.encode_plane_shapes(desc)
rng = np.random.default_rng(1)
frame = rng.random((H, W * 3), dtype=np.float32)


def pinned_torch(shape, dtype):
    return torch.empty(shape, dtype=dtype, pin_memory=True).numpy()


def pinned_lib(shape, dtype):
    count = int(np.prod(shape))
    itemsize = np.dtype(dtype).itemsize
    pointer = C.c_void_p()
    ctx._check(ctx.lib.avifgpu_host_alloc(ctx.handle, count * itemsize, C.byref(pointer)))
    buffer = (C.c_uint8 * (count * itemsize)).from_address(pointer.value)
    return np.frombuffer(buffer, dtype=dtype).reshape(shape)


def planes_struct(arrays):
    p = abi.Planes()
    for k, a in enumerate(arrays):
        if a is not None:
            p.data[k] = a.ctypes.data
            p.stride[k] = a.strides[0]
    return p


def timed(fn):
    fn()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
    return {"ms_mean": 1e3 * sum(times) / len(times), "ms_best": 1e3 * min(times)}


def variant(rows, planes, blocks):
    p = planes_struct(planes)

    def run():
        if blocks == 1:
            ctx._check(ctx.lib.avifgpu_encode_rows(ctx.handle, C.byref(desc), rows.ctypes.data, rows.strides[0], 0, H, C.byref(p)))
            return
        ticket = C.c_int64()
        for top in range(0, H, blocks):
            n = min(blocks, H - top)
            block = rows[top:top + n]
            ctx._check(ctx.lib.avifgpu_encode_rows_async(ctx.handle, C.byref(desc), block.ctypes.data, rows.strides[0], top, n, C.byref(p), C.byref(ticket)))
        ctx._check(ctx.lib.avifgpu_wait(ctx.handle, 0))
    return timed(run)


results = {}
rows_t = pinned_torch(frame.shape, torch.float32)
rows_t[:] = frame
planes_t = [None if s is None else pinned_torch(s, torch.int16) for s in shapes]
rows_l = pinned_lib(frame.shape, np.float32)
rows_l[:] = frame
planes_l = [None if s is None else pinned_lib(s, np.int16) for s in shapes]
planes_pageable = [None if s is None else np.ones(s, dtype=np.int16) for s in shapes]

results["A  one call, torch-pinned rows + planes"] = variant(rows_t, planes_t, 1)
results["A' one call, avifgpu_host_alloc rows + planes"] = variant(rows_l, planes_l, 1)
results["B  6 async calls, lib-pinned rows + planes"] = variant(rows_l, planes_l, 728)
results["C  6 async calls, lib-pinned rows, pageable planes"] = variant(rows_l, planes_pageable, 728)
results["D  one call, lib-pinned rows, pageable planes"] = variant(rows_l, planes_pageable, 1)
results["E  one call, torch-pinned rows, pageable planes"] = variant(rows_t, planes_pageable, 1)
for k, v in results.items():
    print(json.dumps({"variant": k, **v}))
