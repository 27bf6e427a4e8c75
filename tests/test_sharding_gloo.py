"""The N > 1 path on CPU: two gloo ranks shard one frame into row blocks, convert their block (the CPU checker
stands in for the kernel -- this test is about the partition and the gather), all_gather the planes and must
reproduce the single-rank planes exactly, for every chroma mode and awkward heights."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
from avifgpu import abi, sharding


def test_row_blocks_are_even_and_cover():
    for height in (0, 1, 2, 3, 7, 8, 23, 4320, 16384, 4319):
        for parts in (1, 2, 3, 4, 8):
            blocks = sharding.row_blocks(height, parts)
            assert len(blocks) == parts
            assert sum(n for _, n in blocks) == height
            y = 0
            for y0, n in blocks:
                assert y0 == y and n >= 0
                if y0 < height:
                    assert y0 % 2 == 0
                y += n


def test_c_abi_row_blocks_match_the_python_partition():
    """avifgpu_shard_row_blocks (what the in-process shard group cuts by) == sharding.row_blocks (what the ranks cut by)."""
    import avifgpu
    for height in (0, 1, 2, 3, 7, 8, 23, 4320, 16384, 4319):
        for parts in (1, 2, 3, 4, 8):
            assert avifgpu.shard_row_blocks(0, height, parts) == sharding.row_blocks(height, parts)
    # a block of an image (the shuttle's case): boundaries stay on even IMAGE rows
    blocks = avifgpu.shard_row_blocks(512, 1000, 3)
    assert blocks[0][0] == 512 and sum(n for _, n in blocks) == 1000 and all(y0 % 2 == 0 for y0, _ in blocks)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        checker = oracle.load_restatement()
        ok = True
        for (w, h), chroma in [((37, 23), abi.CHROMA_420), ((16, 2), abi.CHROMA_420), ((9, 1), abi.CHROMA_420),
                               ((21, 11), abi.CHROMA_422), ((8, 5), abi.CHROMA_444)]:
            desc = abi.EncodeDesc(w, h, 32, 4, abi.ALPHA_STRAIGHT, 12, abi.TRANSFER_PQ, 80, abi.LAYOUT_PLANAR_YCBCR, chroma,
                                  abi.DOWN_FILTER_BOX, abi.GRAY16_LUT, cases.NCLX_2020_PQ())
            rows = cases.float_host_rows(np.random.default_rng(100 + w), h, w, 4)  # same on both ranks
            expected = checker.encode(desc, rows)
            blocks = sharding.row_blocks(h, world)
            y0, n = blocks[rank]
            shapes = sharding.max_block_plane_shapes(desc, blocks)
            local = [None if s is None else torch.zeros(s, dtype=torch.int16) for s in shapes]
            if n > 0:
                got = checker.encode(sharding.block_desc(desc, n), rows[y0:y0 + n])
                for t, g in zip(local, got):
                    if g is not None:
                        t[:g.shape[0], :g.shape[1]] = torch.from_numpy(g.view(np.int16).copy())
            full = sharding.gather_encode_planes(dist, torch, desc, blocks, local)
            for e, f in zip(expected, full):
                if e is not None:
                    ok = ok and np.array_equal(f.numpy().view(np.uint16), e)
            # gather to the owner only, received in place (what bench.py's tile block times as the NCCL baseline)
            exact = [None if s is None else torch.zeros(s, dtype=torch.int16) for s in abi.encode_plane_shapes(sharding.block_desc(desc, n))]
            for t, l in zip(exact, local):
                if t is not None:
                    t.copy_(l[:t.shape[0], :t.shape[1]])
            owner_planes = [None if e is None else torch.full(e.shape, -1, dtype=torch.int16) for e in expected] if rank == 0 else None
            sharding.gather_planes_to_owner(dist, torch, desc, blocks, exact, owner_planes, rank, owner=0)
            if rank == 0:
                for e, f in zip(expected, owner_planes):
                    if e is not None:
                        ok = ok and np.array_equal(f.numpy().view(np.uint16), e)
        results[rank] = ok
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather():
    world = 2
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    assert dict(results) == {0: True, 1: True}
