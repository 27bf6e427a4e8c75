#!/usr/bin/env python3
"""Two launches (one warm, one to capture) of the tuned float decode kernel for config 3 (HLG + OOTF) and for its PQ sibling,
device-resident, for `ncu -k regex:DecodeYccToRgbF32 --launch-skip ...`.  Prints nothing but the launch count."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "avif-format_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import avifgpu  # noqa: E402
import cases  # noqa: E402
from avifgpu import abi  # noqa: E402

W, H = 7680, 4320
dev = torch.device("cuda", 0)
gpu = avifgpu.Context(0)
g = torch.Generator(device=dev)
g.manual_seed(3)
for nclx, kwargs in ((cases.NCLX_2020_HLG(1), dict(hlg_apply_ootf=1)), (cases.NCLX_2020_PQ(1), dict(pq_peak_nits=80))):
    desc = abi.DecodeDesc(W, H, abi.COLORSPACE_YCBCR, abi.CHROMA_420, 10, abi.ALPHA_NONE, 32, nclx, **kwargs)
    shapes = abi.decode_plane_shapes(desc)
    planes = [None if s is None else torch.randint(0, 1024, s, dtype=torch.int16, device=dev, generator=g) for s in shapes]
    out = torch.empty((H, W * 3), dtype=torch.float32, device=dev)
    struct = avifgpu.planes_from_tensors(planes)
    for _ in range(2):
        gpu.decode_device(desc, struct, out.data_ptr(), out.stride(0) * 4)
    torch.cuda.synchronize()
print(gpu.launch_count())
