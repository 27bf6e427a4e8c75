"""BASELINE.json's full-size configurations on a B200.  The GPU box has enough host cores that the CPU checker
(row-block parallel) converts whole frames in seconds, so configs 2-4 are compared bit for bit at full size; config 5
is compared on two full 4096x4096 images of the batch.  Size-independent properties ride along: row-block additivity
(the N-GPU sharding contract), a round trip through the decoder, idempotence of repeated launches."""
import os

import numpy as np
import pytest

import cases
from avifgpu import abi

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 8


def test_config2_8k_rgb32f_to_12bit_pq_420(gpu, port):
    w, h = 7680, 4320
    desc = abi.EncodeDesc(w, h, 32, 3, abi.ALPHA_NONE, 12, abi.TRANSFER_PQ, 80, abi.LAYOUT_PLANAR_YCBCR, abi.CHROMA_420,
                          abi.DOWN_FILTER_BOX, abi.GRAY16_LUT, cases.NCLX_2020_PQ())
    rows = cases.float_host_rows(np.random.default_rng(2 * 1000 + 1234), h, w, 3)
    expected = port.encode(desc, rows, threads=THREADS)
    got = gpu.encode(desc, rows)
    for k, (e, g) in enumerate(zip(expected, got)):
        if e is not None:
            assert np.array_equal(e, g), f"plane {k}: {int((e != g).sum())} of {e.size} codes differ"
    # the sharding contract: any even row-block partition reproduces the frame
    planes = None
    for y0, n in ((0, 1080), (1080, 2160), (3240, 1080)):
        planes = gpu.encode(desc, rows[y0:y0 + n], y0=y0, nrows=n, planes=planes)
    assert cases.same_planes(expected, planes)
    # idempotence: a second launch over the same buffers changes nothing
    assert cases.same_planes(expected, gpu.encode(desc, rows, planes=got))
    stats = gpu.prepare_encode(desc).as_dict()
    assert stats["valid"] == 1 and stats["verify_mismatches"] == 0


def test_config3_8k_10bit_hlg_420_to_rgb32f(gpu, checker):
    w, h = 7680, 4320
    desc = abi.DecodeDesc(w, h, abi.COLORSPACE_YCBCR, abi.CHROMA_420, 10, abi.ALPHA_NONE, 32, cases.NCLX_2020_HLG(), 1, 1.2, 1000, 80)
    planes = cases.code_planes(np.random.default_rng(3 * 1000 + 1234), desc, overshoot=True)
    expected = checker.decode(desc, planes, threads=THREADS)
    got = gpu.decode(desc, planes)
    assert cases.same_bits(expected, got), f"{int((expected.view(np.uint32) != got.view(np.uint32)).sum())} samples differ"


def test_config4_16k_rgba16_to_10bit_422_alpha(gpu, port):
    w, h = 16384, 16384
    desc = abi.EncodeDesc(w, h, 16, 4, abi.ALPHA_STRAIGHT, 10, abi.TRANSFER_CLIP, 80, abi.LAYOUT_PLANAR_YCBCR, abi.CHROMA_422,
                          abi.DOWN_FILTER_BOX, abi.GRAY16_LUT, None)
    rng = np.random.default_rng(4 * 1000 + 1234)
    rows = rng.integers(0, 32769, (h, w * 4), dtype=np.uint16)
    rows[::97, ::5] = 32768
    expected = port.encode(desc, rows, threads=THREADS)
    got = gpu.encode(desc, rows)
    for k, (e, g) in enumerate(zip(expected, got)):
        if e is not None:
            assert np.array_equal(e, g), f"plane {k} differs"
    # 8-way row tiling (the 8xB200 layout of BASELINE config 4) gives the same planes
    planes = None
    for r in range(8):
        planes = gpu.encode(desc, rows[r * 2048:(r + 1) * 2048], y0=r * 2048, nrows=2048, planes=planes)
    assert cases.same_planes(expected, planes)


def test_config5_gray16_to_12bit_smpte428(gpu, port):
    w, h = 4096, 4096
    rng = np.random.default_rng(5 * 1000 + 1234)
    for image in range(2):
        rows = rng.integers(0, 32769, (h, w), dtype=np.uint16)
        desc = abi.EncodeDesc(w, h, 16, 1, abi.ALPHA_NONE, 12, gray16_curve=abi.GRAY16_SMPTE428)
        assert cases.same_planes(port.encode(desc, rows, threads=THREADS), gpu.encode(desc, rows))
    # a batch is one tall image to the library: 4 images stacked == 4 images converted one by one
    stack = rng.integers(0, 32769, (4 * 512, w), dtype=np.uint16)
    tall = gpu.encode(abi.EncodeDesc(w, 4 * 512, 16, 1, abi.ALPHA_NONE, 12, gray16_curve=abi.GRAY16_SMPTE428), stack)[0]
    for i in range(4):
        one = gpu.encode(abi.EncodeDesc(w, 512, 16, 1, abi.ALPHA_NONE, 12, gray16_curve=abi.GRAY16_SMPTE428), stack[i * 512:(i + 1) * 512])[0]
        assert np.array_equal(tall[i * 512:(i + 1) * 512], one)


def test_config1_512_rgba8_to_yuv444_8bit(gpu, port, checker):
    w = h = 512
    rng = np.random.default_rng(1 * 1000 + 1234)
    rows = cases.int_host_rows(rng, h, w, 4, 8)
    ref_layout = abi.EncodeDesc(w, h, 8, 4, abi.ALPHA_STRAIGHT, 8)
    assert cases.same_planes(checker.encode(ref_layout, rows), gpu.encode(ref_layout, rows))  # the reference's own output
    planar = abi.EncodeDesc(w, h, 8, 4, abi.ALPHA_STRAIGHT, 8, layout=abi.LAYOUT_PLANAR_YCBCR, chroma=abi.CHROMA_444)
    got = gpu.encode(planar, rows)
    assert cases.same_planes(port.encode(planar, rows), got)
    # and back through the reference decoder: within 2 codes (test_forward_roundtrip.py explains the bound)
    back = checker.decode(abi.DecodeDesc(w, h, abi.COLORSPACE_YCBCR, abi.CHROMA_444, 8, abi.ALPHA_STRAIGHT, 8, None), got)
    assert np.abs(back.astype(np.int32) - rows.astype(np.int32)).max() <= 2
