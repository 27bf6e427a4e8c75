"""The host-pointer pipeline beyond the plain call (B200): pageable caller memory on both sides (bounced through pinned
slots), the asynchronous entry points with tickets the double-buffered row shuttle is built on, and the shard group that
puts several GPUs of one process behind the same seam.  Every result must equal the plain single-call conversion bit for
bit -- which test_gpu_parity.py in turn holds against the CPU checker."""
import ctypes as C

import numpy as np
import pytest

import cases
from avifgpu import abi

pytestmark = pytest.mark.gpu


def device_count():
    import torch
    return torch.cuda.device_count()


def c2_desc(w, h, chroma=abi.CHROMA_420):
    return abi.EncodeDesc(w, h, 32, 3, abi.ALPHA_NONE, 12, abi.TRANSFER_PQ, 80, abi.LAYOUT_PLANAR_YCBCR, chroma, abi.DOWN_FILTER_BOX,
                          abi.GRAY16_LUT, cases.NCLX_2020_PQ())


def c4_desc(w, h):
    return abi.EncodeDesc(w, h, 16, 4, abi.ALPHA_STRAIGHT, 10, abi.TRANSFER_CLIP, 80, abi.LAYOUT_PLANAR_YCBCR, abi.CHROMA_422,
                          abi.DOWN_FILTER_BOX, abi.GRAY16_LUT, None)


def hlg_desc(w, h):
    return abi.DecodeDesc(w, h, abi.COLORSPACE_YCBCR, abi.CHROMA_420, 10, abi.ALPHA_NONE, 32, cases.NCLX_2020_HLG(), 1, 1.2, 1000, 80)


def pinned_like(gpu, array):
    out = gpu.pinned_array(array.shape, array.dtype)
    out[...] = array
    return out


def test_pageable_and_pinned_memory_give_the_same_planes(gpu):
    """Large enough for several pipeline slices per call (a slice is ~32 MiB of host rows)."""
    import avifgpu
    w, h = 4096, 2050
    desc = c2_desc(w, h)
    rows = cases.float_host_rows(np.random.default_rng(5), h, w, 3)
    pageable = gpu.encode(desc, rows)  # numpy memory on both sides
    pinned_rows = pinned_like(gpu, rows)
    pinned_planes = [None if s is None else gpu.pinned_array(s, np.uint16) for s in abi.encode_plane_shapes(desc)]
    gpu.encode(desc, pinned_rows, planes=pinned_planes)
    assert cases.same_planes(pageable, pinned_planes)
    # mixed: pinned rows, pageable planes (the plug-in's situation: its own row buffer, libheif's planes)
    mixed = gpu.encode(desc, pinned_rows)
    assert cases.same_planes(pageable, mixed)
    # and the decode direction
    ddesc = hlg_desc(w, h)
    planes = cases.code_planes(np.random.default_rng(6), ddesc)
    expected = gpu.decode(ddesc, planes)
    out = gpu.pinned_array(expected.shape, np.float32)
    gpu.decode(ddesc, [None if p is None else pinned_like(gpu, p) for p in planes], out=out)
    assert cases.same_bits(expected, out)


@pytest.mark.parametrize("block_rows", [2, 64, 510])
def test_async_encode_with_two_row_buffers_equals_the_whole_image(gpu, block_rows):
    """The double-buffered shuttle: block k+1 is produced into the other pinned buffer while block k is on the wire; after
    call k+1 returns, buffer k may be overwritten.  The planes are complete after wait()."""
    w, h = 1920, 1083
    desc = c2_desc(w, h)
    rows = cases.float_host_rows(np.random.default_rng(7), h, w, 3)
    expected = gpu.encode(desc, rows)
    buffers = [gpu.pinned_array((block_rows, w * 3), np.float32) for _ in range(2)]
    planes = [None if s is None else np.zeros(s, np.uint16) for s in abi.encode_plane_shapes(desc)]  # pageable, like libheif's
    tickets = []
    for k, y0 in enumerate(range(0, h, block_rows)):
        n = min(block_rows, h - y0)
        buffer = buffers[k % 2]
        buffer[:n] = rows[y0:y0 + n]          # "advanceState": the host fills the buffer the call before last has released
        tickets.append(gpu.encode_async(desc, buffer[:n], planes, y0=y0, nrows=n))
        buffer_other = buffers[(k + 1) % 2]
        buffer_other[...] = np.nan            # whatever was in the other buffer may be destroyed now
    assert tickets == sorted(tickets) and len(set(tickets)) == len(tickets)
    gpu.wait()
    assert cases.same_planes(expected, planes)


def test_async_decode_tickets_complete_in_order(gpu):
    w, h = 2048, 600
    ddesc = hlg_desc(w, h)
    planes = cases.code_planes(np.random.default_rng(8), ddesc)
    expected = gpu.decode(ddesc, planes)
    block = 150
    outs = [gpu.pinned_array((block, w * 3), np.float32) for _ in range(2)]
    got = np.zeros_like(expected)
    previous = None
    for k, y0 in enumerate(range(0, h, block)):
        out = outs[k % 2]
        ticket = gpu.decode_async(ddesc, planes, out, y0=y0, nrows=block)
        if previous is not None:
            gpu.wait(previous[0])                                      # block k-1 is complete while block k converts
            got[previous[1]:previous[1] + block] = outs[(k - 1) % 2]   # "advanceState": the host takes the rows
        previous = (ticket, y0)
    gpu.wait(previous[0])
    got[previous[1]:previous[1] + block] = outs[(h // block - 1) % 2]
    assert cases.same_bits(expected, got)


def groups():
    n = device_count()
    return [[0]] + ([[0, 1]] if n >= 2 else []) + ([list(range(n))] if n > 2 else [])


@pytest.mark.parametrize("devices", groups(), ids=lambda d: f"{len(d)}gpu")
def test_sharded_host_calls_equal_the_single_gpu_call(gpu, devices):
    import avifgpu
    with avifgpu.ShardGroup(devices) as group:
        assert group.size() == len(devices)
        for desc, rows in ((c2_desc(1000, 1001), cases.float_host_rows(np.random.default_rng(9), 1001, 1000, 3)),
                           (c4_desc(1024, 514), cases.int_host_rows(np.random.default_rng(10), 514, 1024, 4, 16))):
            group.prepare_encode(desc)
            expected = gpu.encode(desc, rows)
            assert cases.same_planes(expected, group.encode(desc, rows))
            # a row block of the image, the way the shuttle presents one
            first, count = 128, (desc.height - 128) // 2 & ~1
            partial = [None if p is None else np.zeros_like(p) for p in expected]
            group.encode(desc, rows[first:first + count], y0=first, nrows=count, planes=partial)
            reference = [None if p is None else np.zeros_like(p) for p in expected]
            gpu.encode(desc, rows[first:first + count], y0=first, nrows=count, planes=reference)
            assert cases.same_planes(reference, partial)
        ddesc = hlg_desc(1000, 1001)
        planes = cases.code_planes(np.random.default_rng(11), ddesc)
        assert cases.same_bits(gpu.decode(ddesc, planes), group.decode(ddesc, planes))
        assert group.launch_count() > 0


@pytest.mark.parametrize("devices", groups(), ids=lambda d: f"{len(d)}gpu")
def test_sharded_device_call_places_every_block_in_the_owners_planes(gpu, devices):
    """Rows distributed over the members' HBM, planes on the owner: the kernels of the other members store across NVLink."""
    import torch
    import avifgpu
    w, h = 2048, 1030
    desc = c4_desc(w, h)
    rows = cases.int_host_rows(np.random.default_rng(12), h, w, 4, 16)
    expected = gpu.encode(desc, rows)
    with avifgpu.ShardGroup(devices) as group:
        n = group.size()
        for member in range(n):
            assert group.peer_access(member, 0), "this box offers no peer access between its GPUs"
        blocks = avifgpu.shard_row_blocks(0, h, n)
        device_rows = []
        for member, (y0, count) in enumerate(blocks):
            device_rows.append(torch.from_numpy(rows[y0:y0 + count].view(np.int16).copy()).to(f"cuda:{devices[member]}"))
        owner = torch.device(f"cuda:{devices[0]}")
        planes = [None if s is None else torch.zeros(s, dtype=torch.int16, device=owner) for s in abi.encode_plane_shapes(desc)]
        group.encode_device(desc, [t.data_ptr() for t in device_rows], [t.stride(0) * 2 for t in device_rows],
                            avifgpu.planes_from_tensors(planes), owner=0)
        group.synchronize()
        got = [None if p is None else p.cpu().numpy().view(np.uint16) for p in planes]
        assert cases.same_planes(expected, got)
