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

