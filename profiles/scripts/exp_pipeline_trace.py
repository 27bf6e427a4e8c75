#!/usr/bin/env python3
"""Experiment helper: one variant of exp_pipeline.py per process (argv[1] = pinned | pageable, argv[2] = rows per call),
so that a library built with pipeline tracing prints its totals for that variant alone when the context is destroyed."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "avif-format_b200", "python"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import avifgpu  # noqa: E402
from avifgpu import abi  # noqa: E402

kind, block = sys.argv[1], int(sys.argv[2])
W, H, steps = 7680, 4320, 10
ctx = avifgpu.Context(0)
nclx = abi.Nclx(1, abi.PRIMARIES_BT2020, abi.TRANSFER_CHAR_PQ, abi.MATRIX_BT2020_NCL, 1)
desc = abi.EncodeDesc(W, H, 32, 3, abi.ALPHA_NONE, 12, abi.TRANSFER_PQ, 80, abi.LAYOUT_PLANAR_YCBCR, abi.CHROMA_420, abi.DOWN_FILTER_BOX, abi.GRAY16_LUT, nclx)
ctx.prepare_encode(desc)
shapes = abi.encode_plane_shapes(desc)
rows = torch.rand((H, W * 3), dtype=torch.float32).pin_memory().numpy()
if kind == "pinned":
    planes = [None if s is None else torch.empty(s, dtype=torch.int16, pin_memory=True).numpy() for s in shapes]
else:
    planes = [None if s is None else np.ones(s, dtype=np.int16) for s in shapes]
p = abi.Planes()
for k, a in enumerate(planes):
    if a is not None:
        p.data[k] = a.ctypes.data
        p.stride[k] = a.strides[0]


def run():
    ticket = C.c_int64()
    for top in range(0, H, block):
        n = min(block, H - top)
        ctx._check(ctx.lib.avifgpu_encode_rows_async(ctx.handle, C.byref(desc), rows[top:top + n].ctypes.data, rows.strides[0], top, n, C.byref(p), C.byref(ticket)))
    ctx._check(ctx.lib.avifgpu_wait(ctx.handle, 0))


run()
t0 = time.perf_counter()
for _ in range(steps):
    run()
print(f"{kind} planes, {block} rows per call: {1e3 * (time.perf_counter() - t0) / steps:.3f} ms per image ({steps + 1} images traced)")
ctx.close()
