"""Row-block sharding of one image across ranks.

Rows are independent (4:2:0: row pairs), so a frame is cut into `world` contiguous row blocks with even boundaries;
every rank converts its block as an image of its own (no halo, no collective on the data path) and the planar
result is assembled with ONE all_gather per plane.  Blocks are padded to the height of the largest block so the
collective is uniform; the padding rows are dropped when the planes are stitched.

Pure torch.distributed plumbing: works with the NCCL backend on GPUs and with gloo on CPU tensors (tests).
"""
from . import abi


def row_blocks(height, parts):
    """[(y0, rows)] * parts: contiguous, covering [0, height), every boundary even (a 2x2 chroma site never
    straddles two blocks); trailing blocks may be empty for tiny images."""
    bounds = [0]
    for i in range(1, parts):
        b = ((height * i) // parts) & ~1
        bounds.append(min(max(b, bounds[-1]), height))
    bounds.append(height)
    return [(bounds[i], bounds[i + 1] - bounds[i]) for i in range(parts)]


def block_desc(desc, rows):
    """The description of a row block presented as an image of `rows` rows."""
    return desc.copy(height=rows)


def max_block_plane_shapes(desc, blocks):
    tallest = max(rows for _, rows in blocks)
    return abi.encode_plane_shapes(block_desc(desc, tallest))


def gather_encode_planes(dist, torch, desc, blocks, local_planes, group=None):
    """local_planes: this rank's block planes (2-D tensors / None), each allocated with the shape of
    max_block_plane_shapes().  Returns the whole-image planes (list of 4 tensors / None) on every rank."""
    world = len(blocks)
    full_shapes = abi.encode_plane_shapes(desc)
    out = []
    for k, local in enumerate(local_planes):
        if local is None:
            out.append(None)
            continue
        gathered = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
        # bytes on the wire: every backend (NCCL, gloo) handles uint8, not all handle int16
        dist.all_gather([gathered[r].view(torch.uint8) for r in range(world)], local.contiguous().view(torch.uint8), group=group)
        rows_total, cols = full_shapes[k]
        full = torch.empty((rows_total, cols), dtype=local.dtype, device=local.device)
        cursor = 0
        for r, (_, rows) in enumerate(blocks):
            shape = abi.encode_plane_shapes(block_desc(desc, rows))[k]
            n = shape[0] if shape is not None else 0
            full[cursor:cursor + n] = gathered[r, :n, :cols]
            cursor += n
        assert cursor == rows_total, (cursor, rows_total)
        out.append(full)
    return out
