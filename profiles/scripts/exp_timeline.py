#!/usr/bin/env python3
"""Experiment (not product): per-CTA start / end times of the config-2 kernel, from a library built with timeline
instrumentation (AVIFGPU_LIBRARY=profiles/scratch_exp/libavifgpu_timeline.so).  Prints where the launch's idle SM time
goes: ramp (launch -> CTA start -> table staged) and tail (first warp done -> last CTA done)."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "avif-format_b200", "python"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import avifgpu  # noqa: E402
from avifgpu import abi  # noqa: E402

W, H = 7680, 4320
dev = torch.device("cuda", 0)
ctx = avifgpu.Context(0)
nclx = abi.Nclx(1, abi.PRIMARIES_BT2020, abi.TRANSFER_CHAR_PQ, abi.MATRIX_BT2020_NCL, 1)
desc = abi.EncodeDesc(W, H, 32, 3, abi.ALPHA_NONE, 12, abi.TRANSFER_PQ, 80, abi.LAYOUT_PLANAR_YCBCR, abi.CHROMA_420, abi.DOWN_FILTER_BOX, abi.GRAY16_LUT, nclx)
ctx.prepare_encode(desc)
shapes = abi.encode_plane_shapes(desc)
g = torch.Generator(device=dev)
g.manual_seed(1)
sets = []
for _ in range(3):
    rows = torch.rand((H, W * 3), generator=g, device=dev)
    planes = [None if s is None else torch.empty(s, dtype=torch.int16, device=dev) for s in shapes]
    sets.append((rows, avifgpu.planes_from_tensors(planes), planes))
lib = ctx.lib if hasattr(ctx, "lib") else avifgpu._lib
fn = lib.avifgpu_debug_timeline
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = np.zeros(148 * 4, dtype=np.uint64)


def launch(i):
    s = sets[i % 3]
    ctx.encode_device(desc, s[0].data_ptr(), s[0].stride(0) * 4, s[1])


for i in range(10):
    launch(i)
torch.cuda.synchronize()
for trial, count in enumerate((1, 1, 4, 4)):
    fn(buf.ctypes.data, 1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(count):
        launch(i)
    b.record()
    torch.cuda.synchronize()
    fn(buf.ctypes.data, 0)
    t = buf.reshape(148, 4).astype(np.int64)
    t0 = t[:, 0].min()
    start = (t[:, 0] - t0) / 1e3
    staged = (t[:, 1] - t[:, 0]) / 1e3
    first_done = (t[:, 2] - t0) / 1e3
    last_done = (t[:, 3] - t0) / 1e3
    print(json.dumps({
        "launches_in_window": count, "event_ms_per_launch": a.elapsed_time(b) / count,
        "note": "times in us relative to the earliest CTA start of the LAST launch in the window",
        "cta_start_us": {"min": float(start.min()), "median": float(np.median(start)), "max": float(start.max())},
        "table_staging_us": {"min": float(staged.min()), "median": float(np.median(staged)), "max": float(staged.max())},
        "first_warp_done_us": {"min": float(first_done.min()), "median": float(np.median(first_done)), "max": float(first_done.max())},
        "cta_done_us": {"min": float(last_done.min()), "p10": float(np.percentile(last_done, 10)), "median": float(np.median(last_done)),
                        "p90": float(np.percentile(last_done, 90)), "max": float(last_done.max())},
        "mean_idle_at_tail_us": float((last_done.max() - last_done).mean()),
        "cta_done_sorted_us": [round(float(x), 1) for x in np.sort(last_done)[::8]],
    }))
    if trial == 0:
        order = np.argsort(last_done)
        print(json.dumps({"fastest_ctas": [int(x) for x in order[:16]], "slowest_ctas": [int(x) for x in order[-16:]],
                          "done_by_cta_block_of_8": [round(float(last_done[i:i + 8].mean()), 1) for i in range(0, 148, 8)]}))
