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


def gather_planes_to_owner(dist, torch, desc, blocks, local_planes, owner_planes, rank, owner=0, group=None):
    """The collective of SURVEY.md 8(e) as it should be: every rank sends each plane block ONCE, to the owner only, and
    the owner receives it straight into its rows of the final plane (plane rows are contiguous, so a row block is a
    contiguous slice: no re-stitch copy).  All sends and receives of a call form one group
    (torch.distributed.batch_isend_irecv = ncclGroupStart / ncclGroupEnd on NCCL; works on gloo too).

    local_planes   this rank's block planes, exactly the shape of the block presented as an image
                   (abi.encode_plane_shapes(block_desc(desc, rows)))
    owner_planes   whole-image planes on the owner (ignored elsewhere; may be None)"""
    ys = [0, 1 if desc.chroma == abi.CHROMA_420 and desc.layout == abi.LAYOUT_PLANAR_YCBCR else 0,
          1 if desc.chroma == abi.CHROMA_420 and desc.layout == abi.LAYOUT_PLANAR_YCBCR else 0, 0]
    ops = []
    for k, plane in enumerate(local_planes):
        if plane is None:
            continue
        if rank == owner:
            for r, (y0, n) in enumerate(blocks):
                rows = (n + ys[k]) >> ys[k]
                if rows == 0:
                    continue
                target = owner_planes[k][(y0 >> ys[k]):(y0 >> ys[k]) + rows]
                if r == owner:
                    target.copy_(plane)
                else:
                    ops.append(dist.P2POp(dist.irecv, target.view(torch.uint8), r, group))
        elif plane.numel() > 0:
            ops.append(dist.P2POp(dist.isend, plane.contiguous().view(torch.uint8), owner, group))
    if ops:
        for work in dist.batch_isend_irecv(ops):
            work.wait()


class PeerPlanes:
    """Whole-image planes that live on the OWNER rank's GPU and are mapped into every other rank's address space
    through CUDA IPC (one process per GPU).  A rank hands `planes()` to avifgpu_encode_rows_device together with
    its own (y0, nrows): the conversion kernel's stores then land in the owner's HBM over NVLink / NVSwitch, tile
    by tile, while the kernel is still converting -- the "gather" is fused into the conversion and costs no extra
    pass (the planes are 3 of the 15 bytes per pixel the kernel moves).  After the ranks' streams have drained and a
    barrier, the owner holds the assembled image.

    Needs a GPU per rank and peer access between them (any NVSwitch box); there is no CPU equivalent, so the gloo
    tests cover the all_gather path (gather_encode_planes) and this class is exercised by `bench.py --mode tile`."""

    def __init__(self, dist, desc, rank, owner=0, group=None):
        from cuda.bindings import runtime as rt
        self._rt = rt
        self.rank, self.owner = rank, owner
        self.shapes = abi.encode_plane_shapes(desc)
        self.itemsize = 2 if desc.image_bit_depth > 8 else 1
        self.pointers, self.strides, self._opened = [0] * 4, [0] * 4, []
        meta = [None]
        if rank == owner:
            entries = []
            for shape in self.shapes:
                if shape is None:
                    entries.append(None)
                    continue
                pitch = (shape[1] * self.itemsize + 255) // 256 * 256
                ptr = self._ok(rt.cudaMalloc(pitch * max(shape[0], 1)))
                handle = self._ok(rt.cudaIpcGetMemHandle(ptr))
                entries.append((bytes(handle.reserved), pitch, int(ptr)))
            meta = [entries]
        dist.broadcast_object_list(meta, src=owner, group=group)
        for k, entry in enumerate(meta[0]):
            if entry is None:
                continue
            reserved, pitch, owner_ptr = entry
            if rank == owner:
                ptr = owner_ptr
            else:
                handle = rt.cudaIpcMemHandle_t()
                handle.reserved = reserved
                ptr = int(self._ok(rt.cudaIpcOpenMemHandle(handle, rt.cudaIpcMemLazyEnablePeerAccess)))
                self._opened.append(ptr)
            self.pointers[k], self.strides[k] = ptr, pitch

    def _ok(self, result):
        err = result[0]
        if int(err) != 0:
            raise RuntimeError(f"CUDA runtime error {err} in PeerPlanes")
        return result[1] if len(result) > 1 else None

    def planes(self):
        """abi.Planes over the whole image (origin pointers), valid on this rank's device."""
        planes = abi.Planes()
        for k in range(4):
            planes.data[k] = self.pointers[k] or None
            planes.stride[k] = self.strides[k]
        return planes

    def owner_tensors(self, torch, device):
        """The owner's view of the assembled planes as torch tensors (copies)."""
        assert self.rank == self.owner
        out = []
        for k, shape in enumerate(self.shapes):
            if shape is None:
                out.append(None)
                continue
            dtype = torch.int16 if self.itemsize == 2 else torch.uint8
            pitched = torch.empty((shape[0], self.strides[k] // self.itemsize), dtype=dtype, device=device)
            self._ok(self._rt.cudaMemcpy(pitched.data_ptr(), self.pointers[k], self.strides[k] * shape[0],
                                         self._rt.cudaMemcpyKind.cudaMemcpyDeviceToDevice))
            out.append(pitched[:, :shape[1]].contiguous())
        return out

    def close(self):
        for ptr in self._opened:
            self._rt.cudaIpcCloseMemHandle(ptr)
        self._opened = []
        if self.rank == self.owner:
            for ptr in self.pointers:
                if ptr:
                    self._rt.cudaFree(ptr)
        self.pointers = [0] * 4
