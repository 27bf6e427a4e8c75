"""Seeded case generators shared by the parity tests, the golden-vector script and smoke().

Every case is (name, desc, inputs): small enough that the CPU checkers finish instantly, shaped to hit the edge
cases the reference's loops have: odd widths / heights (chroma edges), 1x1, empty images, alpha 0 / max,
out-of-range codes in 16-bit containers (clamped by YuvDecode.cpp), HDR over-range and negative floats.
"""
import itertools
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (os.path.join(_ROOT, "avif-format_b200", "python"), os.path.join(_ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from avifgpu import abi  # noqa: E402

NCLX_NONE = None
NCLX_601 = lambda full=1: abi.Nclx(1, abi.PRIMARIES_BT709, abi.TRANSFER_CHAR_SRGB, abi.MATRIX_BT601, full)  # noqa: E731
NCLX_709 = lambda full=1: abi.Nclx(1, abi.PRIMARIES_BT709, abi.TRANSFER_CHAR_SRGB, abi.MATRIX_BT709, full)  # noqa: E731
NCLX_2020_PQ = lambda full=1: abi.Nclx(1, abi.PRIMARIES_BT2020, abi.TRANSFER_CHAR_PQ, abi.MATRIX_BT2020_NCL, full)  # noqa: E731
NCLX_2020_HLG = lambda full=1: abi.Nclx(1, abi.PRIMARIES_BT2020, abi.TRANSFER_CHAR_HLG, abi.MATRIX_BT2020_NCL, full)  # noqa: E731
NCLX_2020_428 = lambda full=1: abi.Nclx(1, abi.PRIMARIES_BT2020, abi.TRANSFER_CHAR_SMPTE428, abi.MATRIX_BT2020_NCL, full)  # noqa: E731
NCLX_GBR = lambda: abi.Nclx(1, abi.PRIMARIES_BT709, abi.TRANSFER_CHAR_SRGB, abi.MATRIX_GBR, 1)  # noqa: E731
NCLX_DERIVED = lambda prim=abi.PRIMARIES_BT2020: abi.Nclx(1, prim, abi.TRANSFER_CHAR_PQ, abi.MATRIX_CHROMA_DERIVED_NCL, 1)  # noqa: E731

SIZES = [(37, 23), (64, 16), (1, 1), (2, 2), (3, 1), (1, 4), (130, 5)]


def rng_for(name):
    seed = int.from_bytes(name.encode("utf-8"), "little") % (2 ** 32)
    return np.random.default_rng(seed)


# ---- encode inputs ---------------------------------------------------------------------------------------------

def float_host_rows(rng, h, w, channels, specials=False):
    """SURVEY 8d mix: 70 % uniform[0,1], 20 % log-uniform[1e-4,4], 5 % negatives, 5 % exact {0,1}."""
    n = h * w * channels
    kind = rng.random(n)
    v = rng.random(n, dtype=np.float32)
    logu = np.exp(rng.uniform(np.log(1e-4), np.log(4.0), n)).astype(np.float32)
    v = np.where(kind < 0.20, logu, v)
    v = np.where((kind >= 0.20) & (kind < 0.25), -rng.random(n, dtype=np.float32), v)
    v = np.where((kind >= 0.25) & (kind < 0.30), np.round(rng.random(n)).astype(np.float32), v)
    v = v.astype(np.float32).reshape(h, w, channels)
    if channels in (2, 4):
        a = v[..., -1]
        a[rng.random(a.shape) < 0.15] = 0.0
        a[rng.random(a.shape) < 0.15] = 1.0
        a[rng.random(a.shape) < 0.05] = 1.5
        a[rng.random(a.shape) < 0.05] = -0.25
    if specials and n >= 8:
        flat = v.reshape(-1)
        picks = [np.nan, np.inf, -np.inf, -0.0, 1e-45, 1e-39, 3.4e38, 125.0, 125.5, 10000.0 / 80.0, 1.0000001, 0.99999994]
        idx = rng.choice(n, size=min(len(picks), n), replace=False)
        for i, value in zip(idx, picks):
            flat[i] = value
    return np.ascontiguousarray(v.reshape(h, w * channels))


def int_host_rows(rng, h, w, channels, host_depth, beyond=False):
    top = 255 if host_depth == 8 else 32768
    dtype = np.uint8 if host_depth == 8 else np.uint16
    v = rng.integers(0, top + 1, (h, w, channels)).astype(dtype)
    edge = rng.random((h, w, channels))
    v[edge < 0.05] = 0
    v[edge > 0.95] = top
    if channels in (2, 4):
        a = v[..., -1]
        pick = rng.random(a.shape)
        a[pick < 0.2] = 0
        a[pick > 0.8] = top
    if beyond and host_depth == 16:
        pick = rng.random((h, w, channels)) < 0.05
        v[pick] = rng.integers(32769, 65536, int(pick.sum())).astype(dtype)
    return np.ascontiguousarray(v.reshape(h, w * channels))


def encode_cases(sizes=None, full=True):
    """Yields (name, desc, rows, reference_ok).  reference_ok is False where the compiled reference has undefined
    behaviour or no such path (then only the restatement is the checker)."""
    sizes = sizes or SIZES[:3]
    for (w, h) in sizes:
        # -- integer hosts, reference layout
        for host_depth, channels, depth in itertools.product((8, 16), (1, 2, 3, 4), (8, 10, 12)):
            alphas = [abi.ALPHA_NONE] if channels in (1, 3) else [abi.ALPHA_STRAIGHT, abi.ALPHA_PREMULTIPLIED]
            for alpha in alphas:
                name = f"enc_ref_h{host_depth}_c{channels}_a{alpha}_d{depth}_{w}x{h}"
                rng = rng_for(name)
                desc = abi.EncodeDesc(w, h, host_depth, channels, alpha, depth)
                yield name, desc, int_host_rows(rng, h, w, channels, host_depth), True
        # -- float hosts, reference layout
        for channels, depth, (transfer, peak) in itertools.product(
                (1, 2, 3, 4), (10, 12),
                ((abi.TRANSFER_PQ, 80), (abi.TRANSFER_PQ, 1000), (abi.TRANSFER_PQ, 10000), (abi.TRANSFER_SMPTE428, 80),
                 (abi.TRANSFER_CLIP, 80))):
            if channels <= 2 and transfer == abi.TRANSFER_SMPTE428:
                continue
            alphas = [abi.ALPHA_NONE] if channels in (1, 3) else [abi.ALPHA_STRAIGHT, abi.ALPHA_PREMULTIPLIED]
            for alpha in alphas:
                name = f"enc_ref_h32_c{channels}_a{alpha}_d{depth}_t{transfer}_p{peak}_{w}x{h}"
                rng = rng_for(name)
                desc = abi.EncodeDesc(w, h, 32, channels, alpha, depth, transfer, peak)
                yield name, desc, float_host_rows(rng, h, w, channels), True
        if not full:
            continue
        # -- planar YCbCr (forward matrix + down-filter: this project's definition, restatement only)
        for host_depth, channels, chroma, down, nclx_name in itertools.product(
                (8, 16, 32), (3, 4), (abi.CHROMA_444, abi.CHROMA_422, abi.CHROMA_420),
                (abi.DOWN_FILTER_BOX, abi.DOWN_FILTER_TOP_LEFT), ("none", "709", "2020", "derived", "gbr")):
            if nclx_name == "gbr" and chroma != abi.CHROMA_444:
                continue
            if down == abi.DOWN_FILTER_TOP_LEFT and nclx_name not in ("none", "2020"):
                continue
            nclx = {"none": None, "709": NCLX_709(), "2020": NCLX_2020_PQ(), "derived": NCLX_DERIVED(), "gbr": NCLX_GBR()}[nclx_name]
            depth = {8: 8, 16: 10, 32: 12}[host_depth]
            alpha = abi.ALPHA_NONE if channels == 3 else abi.ALPHA_PREMULTIPLIED if host_depth != 32 else abi.ALPHA_STRAIGHT
            transfer = abi.TRANSFER_PQ if host_depth == 32 else abi.TRANSFER_CLIP
            name = f"enc_ycc_h{host_depth}_c{channels}_ch{chroma}_f{down}_{nclx_name}_{w}x{h}"
            rng = rng_for(name)
            desc = abi.EncodeDesc(w, h, host_depth, channels, alpha, depth, transfer, 80, abi.LAYOUT_PLANAR_YCBCR, chroma, down,
                                  abi.GRAY16_LUT, nclx)
            rows = float_host_rows(rng, h, w, channels) if host_depth == 32 else int_host_rows(rng, h, w, channels, host_depth)
            yield name, desc, rows, False
        # -- HLG save path (the reference has LinearToHLG / ApplyInverseHLGOOTF but no caller: restatement only)
        for channels, extension, layout, depth in itertools.product((3, 4), (abi.HLG_OETF, abi.HLG_INVERSE_OOTF_THEN_OETF),
                                                                    (abi.LAYOUT_REFERENCE, abi.LAYOUT_PLANAR_YCBCR), (10, 12)):
            alphas = [abi.ALPHA_NONE] if channels == 3 else [abi.ALPHA_STRAIGHT, abi.ALPHA_PREMULTIPLIED]
            for alpha in alphas:
                name = f"enc_hlg_c{channels}_a{alpha}_x{extension}_l{layout}_d{depth}_{w}x{h}"
                rng = rng_for(name)
                desc = abi.EncodeDesc(w, h, 32, channels, alpha, depth, abi.TRANSFER_HLG, 80, layout, abi.CHROMA_420, abi.DOWN_FILTER_BOX,
                                      abi.GRAY16_LUT, NCLX_2020_HLG(), hlg_extension=extension, hlg_display_gamma=1.2, hlg_peak_nits=1000)
                yield name, desc, float_host_rows(rng, h, w, channels), False
        # -- colour-profile matrix ahead of the float pipeline (lcms2 is not in the tree: restatement only)
        for channels, alpha, layout, transfer in ((3, abi.ALPHA_NONE, abi.LAYOUT_PLANAR_YCBCR, abi.TRANSFER_PQ), (4, abi.ALPHA_PREMULTIPLIED, abi.LAYOUT_PLANAR_YCBCR, abi.TRANSFER_PQ),
                                                  (4, abi.ALPHA_STRAIGHT, abi.LAYOUT_REFERENCE, abi.TRANSFER_SMPTE428), (3, abi.ALPHA_NONE, abi.LAYOUT_REFERENCE, abi.TRANSFER_CLIP)):
            name = f"enc_rowmatrix_c{channels}_a{alpha}_l{layout}_t{transfer}_{w}x{h}"
            desc = abi.EncodeDesc(w, h, 32, channels, alpha, 12, transfer, 1000, layout, abi.CHROMA_420, abi.DOWN_FILTER_BOX, abi.GRAY16_LUT, NCLX_2020_PQ())
            desc.row_matrix_enabled = 1
            desc.row_matrix = type(desc.row_matrix)(*ROW_MATRIX_709_TO_2020)
            yield name, desc, float_host_rows(rng_for(name), h, w, channels), False
        # -- gray16 -> SMPTE 428 (BASELINE config 5, this project's composition)
        for channels in (1, 2):
            alpha = abi.ALPHA_NONE if channels == 1 else abi.ALPHA_STRAIGHT
            name = f"enc_gray16_428_c{channels}_{w}x{h}"
            rng = rng_for(name)
            desc = abi.EncodeDesc(w, h, 16, channels, alpha, 12, gray16_curve=abi.GRAY16_SMPTE428)
            yield name, desc, int_host_rows(rng, h, w, channels, 16), False
        # -- defined-by-us inputs: specials in float hosts (NaN/inf: the reference's cast is undefined), >32768 samples
        name = f"enc_specials_h32_{w}x{h}"
        yield name, abi.EncodeDesc(w, h, 32, 3, abi.ALPHA_NONE, 12, abi.TRANSFER_PQ, 80), float_host_rows(rng_for(name), h, w, 3, True), False
        name = f"enc_beyond_h16_{w}x{h}"
        yield name, abi.EncodeDesc(w, h, 16, 4, abi.ALPHA_PREMULTIPLIED, 10), int_host_rows(rng_for(name), h, w, 4, 16, True), False


# linear Rec.709 -> linear Rec.2020 (ITU-R BT.2087-0), the matrix a linear sRGB-primaries document needs for an HDR save
ROW_MATRIX_709_TO_2020 = (0.6274039, 0.3292830, 0.0433131, 0.0690973, 0.9195404, 0.0113623, 0.0163914, 0.0880133, 0.8955953)


# ---- decode inputs ---------------------------------------------------------------------------------------------

def code_planes(rng, desc, overshoot=False):
    shapes = abi.decode_plane_shapes(desc)
    dtype = abi.code_dtype(desc.bit_depth)
    top = (1 << desc.bit_depth) - 1
    planes = []
    for i, shape in enumerate(shapes):
        if shape is None:
            planes.append(None)
            continue
        p = rng.integers(0, top + 1, shape).astype(dtype)
        edge = rng.random(shape)
        p[edge < 0.04] = 0
        p[edge > 0.96] = top
        if i == 3:
            pick = rng.random(shape)
            p[pick < 0.2] = 0
            p[pick > 0.8] = top
        elif i in (1, 2) and desc.colorspace == abi.COLORSPACE_YCBCR:
            p[rng.random(shape) < 0.1] = (top + 1) // 2
        if overshoot and desc.bit_depth in (10, 12) and p.size:
            pick = rng.random(shape) < 0.05
            p[pick] = rng.integers(top + 1, 65536, int(pick.sum())).astype(dtype)
        planes.append(np.ascontiguousarray(p))
    return planes


def decode_cases(sizes=None, full=True):
    """Yields (name, desc, planes, reference_ok)."""
    sizes = sizes or SIZES[:3]
    for (w, h) in sizes:
        for chroma, alpha, nclx_name in itertools.product(
                (abi.CHROMA_444, abi.CHROMA_422, abi.CHROMA_420),
                (abi.ALPHA_NONE, abi.ALPHA_STRAIGHT, abi.ALPHA_PREMULTIPLIED), ("none", "709full", "601lim", "2020lim")):
            nclx = {"none": None, "709full": NCLX_709(1), "601lim": NCLX_601(0), "2020lim": NCLX_2020_PQ(0)}[nclx_name]
            for bit_depth, host_depth in ((8, 8), (10, 16), (12, 16), (16, 16)):
                if bit_depth == 16 and not (nclx_name in ("none", "709full")):
                    continue  # 16-bit limited range overflows int in the reference (kept out of the reference gate)
                name = f"dec_ycc_b{bit_depth}_h{host_depth}_ch{chroma}_a{alpha}_{nclx_name}_{w}x{h}"
                desc = abi.DecodeDesc(w, h, abi.COLORSPACE_YCBCR, chroma, bit_depth, alpha, host_depth, nclx)
                yield name, desc, code_planes(rng_for(name), desc, overshoot=True), True
        # YCbCr 4:4:4 carrying the identity (GBR) matrix -- what a lossless file decodes as when libheif hands the planes over
        # as YCbCr: the tables treat chroma like luma (YuvLookupTables.cpp:145,177-180), the row decoders have no identity
        # branch and apply the default coefficients; whatever the reference does with it is the contract
        for alpha, (bit_depth, host_depth) in itertools.product((abi.ALPHA_NONE, abi.ALPHA_STRAIGHT), ((8, 8), (10, 16), (12, 32))):
            name = f"dec_ycc_gbr_b{bit_depth}_h{host_depth}_a{alpha}_{w}x{h}"
            nclx = NCLX_GBR()
            if host_depth == 32:
                nclx = NCLX_2020_PQ()
                nclx.matrix_coefficients = abi.MATRIX_GBR
            desc = abi.DecodeDesc(w, h, abi.COLORSPACE_YCBCR, abi.CHROMA_444, bit_depth, alpha, host_depth, nclx)
            yield name, desc, code_planes(rng_for(name), desc, overshoot=False), True
        # float outputs: PQ / HLG (+OOTF on/off) / SMPTE 428
        for chroma, alpha, (nclx_fn, tname), bit_depth in itertools.product(
                (abi.CHROMA_444, abi.CHROMA_420), (abi.ALPHA_NONE, abi.ALPHA_PREMULTIPLIED),
                ((NCLX_2020_PQ, "pq"), (NCLX_2020_HLG, "hlg"), (NCLX_2020_428, "428")), (10, 12)):
            for variant in range(3):
                name = f"dec_ycc32_b{bit_depth}_ch{chroma}_a{alpha}_{tname}_v{variant}_{w}x{h}"
                full_range = 0 if variant == 2 else 1
                desc = abi.DecodeDesc(w, h, abi.COLORSPACE_YCBCR, chroma, bit_depth, alpha, 32, nclx_fn(full_range),
                                      hlg_apply_ootf=int(variant != 1), hlg_display_gamma=(1.2, 1.2, 1.4)[variant],
                                      hlg_peak_nits=(1000, 1000, 400)[variant], pq_peak_nits=(80, 1000, 10000)[variant])
                yield name, desc, code_planes(rng_for(name), desc, overshoot=True), True
        if not full:
            continue
        # monochrome
        for alpha, (bit_depth, host_depth), full_range in itertools.product(
                (abi.ALPHA_NONE, abi.ALPHA_STRAIGHT, abi.ALPHA_PREMULTIPLIED), ((8, 8), (10, 16), (12, 16), (10, 32), (12, 32)), (1, 0)):
            name = f"dec_mono_b{bit_depth}_h{host_depth}_a{alpha}_r{full_range}_{w}x{h}"
            nclx = NCLX_2020_PQ(full_range)
            desc = abi.DecodeDesc(w, h, abi.COLORSPACE_MONOCHROME, abi.CHROMA_MONOCHROME, bit_depth, alpha, host_depth, nclx,
                                  pq_peak_nits=80 if full_range else 1000)
            yield name, desc, code_planes(rng_for(name), desc, overshoot=True), True
        # planar RGB
        for alpha, (bit_depth, host_depth) in itertools.product(
                (abi.ALPHA_NONE, abi.ALPHA_STRAIGHT, abi.ALPHA_PREMULTIPLIED), ((8, 8), (10, 16), (12, 16))):
            name = f"dec_rgb_b{bit_depth}_h{host_depth}_a{alpha}_{w}x{h}"
            desc = abi.DecodeDesc(w, h, abi.COLORSPACE_RGB, abi.CHROMA_444, bit_depth, alpha, host_depth, NCLX_GBR())
            yield name, desc, code_planes(rng_for(name), desc, overshoot=(host_depth == 16)), True
        for alpha, bit_depth, (nclx_fn, tname) in itertools.product(
                (abi.ALPHA_NONE, abi.ALPHA_PREMULTIPLIED), (10, 12), ((NCLX_2020_PQ, "pq"), (NCLX_2020_HLG, "hlg"), (NCLX_2020_428, "428"))):
            name = f"dec_rgb32_b{bit_depth}_a{alpha}_{tname}_{w}x{h}"
            nclx = nclx_fn()
            nclx.matrix_coefficients = abi.MATRIX_GBR
            desc = abi.DecodeDesc(w, h, abi.COLORSPACE_RGB, abi.CHROMA_444, bit_depth, alpha, 32, nclx)
            yield name, desc, code_planes(rng_for(name), desc, overshoot=False), True


def same_planes(a, b):
    for pa, pb in zip(a, b):
        if (pa is None) != (pb is None):
            return False
        if pa is not None and not np.array_equal(pa, pb):
            return False
    return True


def same_bits(a, b):
    """Bit-for-bit equality (float arrays compared as integers so NaN payloads and signed zeros count)."""
    if a.dtype == np.float32:
        return np.array_equal(a.view(np.uint32), b.view(np.uint32))
    return np.array_equal(a, b)
