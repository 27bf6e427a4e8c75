"""Property tests of the host-side logic that needs no device: row-block partitioning (the multi-GPU sharding contract),
plane geometry, and the byte accounting bench.py's roofline is computed from."""
import os
import sys

import pytest
from hypothesis import given, settings, strategies as st

from avifgpu import abi, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@settings(max_examples=300, deadline=None)
@given(height=st.integers(0, 70000), parts=st.integers(1, 16))
def test_row_blocks_partition_the_image_on_even_boundaries(height, parts):
    blocks = sharding.row_blocks(height, parts)
    assert len(blocks) == parts
    cursor = 0
    for index, (y0, rows) in enumerate(blocks):
        assert y0 == cursor and rows >= 0
        if index > 0:
            assert y0 % 2 == 0 or y0 == height  # a 2x2 chroma site never straddles two blocks
        cursor += rows
    assert cursor == height
    sizes = [rows for _, rows in blocks if rows]
    if height >= 2 * parts:
        assert max(sizes) - min(sizes) <= 2 + (height % 2)  # balanced up to the even rounding


@settings(max_examples=200, deadline=None)
@given(w=st.integers(1, 9000), h=st.integers(1, 9000), chroma=st.sampled_from([abi.CHROMA_444, abi.CHROMA_422, abi.CHROMA_420]),
       alpha=st.booleans(), depth=st.sampled_from([8, 10, 12]))
def test_encode_plane_shapes_follow_the_chroma_format(w, h, chroma, alpha, depth):
    host_depth = 8 if depth == 8 else 16
    desc = abi.EncodeDesc(w, h, host_depth, 4 if alpha else 3, abi.ALPHA_STRAIGHT if alpha else abi.ALPHA_NONE, depth, abi.TRANSFER_CLIP, 80,
                          abi.LAYOUT_PLANAR_YCBCR, chroma)
    shapes = abi.encode_plane_shapes(desc)
    xs, ys = abi.chroma_shifts(chroma)
    assert shapes[0] == (h, w)
    assert shapes[1] == shapes[2] == ((h + ys) >> ys, (w + xs) >> xs)
    assert (shapes[3] == (h, w)) if alpha else (shapes[3] is None)
    # a block's planes are the image's planes restricted to the block (what avifgpu.sharding relies on)
    for y0, rows in sharding.row_blocks(h, 3):
        if rows:
            block = abi.encode_plane_shapes(sharding.block_desc(desc, rows))
            assert block[0] == (rows, w) and block[1][1] == shapes[1][1]


def test_bench_byte_accounting_matches_the_survey():
    sys.path.insert(0, ROOT)
    import bench
    expected = {"c2": (33177600, 497664000), "c3": (33177600, 497664000), "c4": (268435456, 3758096384), "c5": (536870912, 2147483648)}
    for key, (pixels, algorithmic_bytes) in expected.items():
        wl = bench.Workload(key)
        assert wl.pixels == pixels, key
        assert wl.algorithmic_bytes == algorithmic_bytes, key


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference runs without a GPU (it is the CPU arm) and prints ONE JSON line with the keys the driver reads,
    the same `config` object as the GPU arm's, and a cpu_baseline that says which library ran which stage."""
    import json
    import subprocess
    env = dict(os.environ, AVIFGPU_BENCH_REFERENCE_BUDGET_S="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [line for line in out.stdout.splitlines() if line.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "Gpx/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["e2e"] == {"value": line["value"], "unit": "Gpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.common_config(bench.Workload("c2"))
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1 and "stages" in line["cpu_baseline"]



# ---- the premise of the float decode kernel's branch-free powf --------------------------------------------------------------

@pytest.mark.parametrize("bit_depth", [10, 12])
@pytest.mark.parametrize("nclx_name", ["none", "601lim", "709full", "2020full", "2020lim", "derived2020", "gbr"])
def test_clamped_channel_sums_are_zero_or_far_from_subnormal(bit_depth, nclx_name):
    """kernels_fast_decode.cu raises the clamped channel sums R = Y + rGain Cr, B = Y + bGain Cb, G = Y - gTerm(Cb, Cr) to PQ's
    1 / m2 and SMPTE 428's 2.6 with a powf that takes +0 or NORMAL bases and a log2 table that starts at 2^-96
    (ChannelSumsStayNormal, LowestWideExponent).  With the CPU checker's own tables and coefficients (YuvLookupTables.cpp:157-184,
    YUVCoefficiants.cpp), in the reference's float32 expressions (YuvDecode.cpp:306-312): over EVERY (Y, Cr) and (Y, Cb) code pair,
    and for G over every Y against the green terms nearest to it (the only candidates for a tiny difference), the smallest
    non-zero clamped sum stays above 2^-40."""
    import numpy as np
    import cases
    import oracle

    nclx = {"none": None, "601lim": cases.NCLX_601(0), "709full": cases.NCLX_709(1), "2020full": cases.NCLX_2020_PQ(1),
            "2020lim": cases.NCLX_2020_PQ(0), "derived2020": cases.NCLX_DERIVED(), "gbr": cases.NCLX_GBR()}[nclx_name]
    checker = oracle.load_restatement()
    kr, kg, kb = (np.float32(v) for v in checker.yuv_coefficients(nclx))
    table_y, table_uv, _ = checker.yuv_tables(nclx, bit_depth, False)
    one, two = np.float32(1), np.float32(2)
    r_gain = two * (one - kr)
    b_gain = two * (one - kb)
    # the launcher's own screen (ChannelSumsStayNormal): every H.273 matrix passes it
    for factor in (r_gain, b_gain, kr * (one - kr), kb * (one - kb), kg):
        assert np.float32(1 / 65536) <= factor <= np.float32(4)

    def smallest_positive(values):
        clamped = np.clip(values, np.float32(0), np.float32(1))
        positive = clamped[clamped > 0]
        return float(positive.min()) if positive.size else 1.0

    least = 1.0
    for gain in (r_gain, b_gain):
        offsets = (gain * table_uv).astype(np.float32)               # one rounding, as in the kernel and the reference
        sums = (table_y[:, None] + offsets[None, :]).astype(np.float32)
        least = min(least, smallest_positive(sums))
    # G = Y - (2 (kr (1 - kr) Cr + kb (1 - kb) Cb)) / kg: all 2^(2 depth) green terms, then per Y the terms nearest to it
    g_cr = ((kr * (one - kr)) * table_uv).astype(np.float32)
    g_cb = ((kb * (one - kb)) * table_uv).astype(np.float32)
    numerators = (two * (g_cr[:, None] + g_cb[None, :]).astype(np.float32)).astype(np.float32)
    green_terms = np.unique((numerators / kg).astype(np.float32))
    for y in np.unique(table_y):
        at = int(np.searchsorted(green_terms, y))
        near = green_terms[max(0, at - 4):at + 4]
        least = min(least, smallest_positive((y - near).astype(np.float32)))
    assert least > 2.0 ** -40, (nclx_name, bit_depth, least)


def test_pq_quotient_is_zero_or_far_from_subnormal():
    """The second powf of PQToLinear (ColorTransfer.cpp:110-112) takes max(x - c1, 0) / (c2 - c3 x) with x = powf(value, 1 / m2) in
    [0, 1]; the tuned decode kernel evaluates it on +0 or NORMAL bases inside a log2 table that starts at 2^-96.  Over every float
    x in [c1, 1] (below c1 the numerator is +0), in the reference's float32 expressions: the quotient is 0, or at least 2^-29, and
    never above 1."""
    import numpy as np
    c1 = np.float32(3424.0) / np.float32(4096.0)
    c2 = np.float32(2413.0) / np.float32(4096.0) * np.float32(32.0)
    c3 = np.float32(2392.0) / np.float32(4096.0) * np.float32(32.0)
    first = int(np.array([c1], np.float32).view(np.uint32)[0]) - 8
    x = np.arange(first, 0x3F800001, dtype=np.uint32).view(np.float32)
    numerator = np.maximum((x - c1).astype(np.float32), np.float32(0))
    denominator = (c2 - (c3 * x).astype(np.float32)).astype(np.float32)
    assert float(denominator.min()) > 0.16 and float(denominator.max()) < 18.9
    quotient = (numerator / denominator).astype(np.float32)
    assert float(quotient.max()) <= 1.0
    positive = quotient[quotient > 0]
    assert float(positive.min()) >= 2.0 ** -29
    assert not bool((quotient[:8] != 0).any())  # x < c1
