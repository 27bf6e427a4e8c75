"""Pins this project's definition of the libheif stage (forward matrix + quantisation) two ways.

1. Against the STANDARD, independently of any code of this repo: H.273's full-range equations evaluated here in float64
   (E'Y = kr R + kg G + kb B, E'Cb = (B - E'Y) / (2 (1 - kb)), Ycode = Clip(Round(E'Y)), Ccode = Clip(Round(E'C) +
   2^(depth-1))) must agree with the float32 implementation on every sample up to the one code a float32 rounding can move
   a value sitting on a .5 boundary; and the H.273 inverse must return the RGB codes within the quantisation bound.
2. By a round trip through the reference's own decoder (the compiled DecodeYUV*Row*).  That decoder puts the chroma zero
   at max/2 (YuvLookupTables.cpp:183: i/max - 0.5f), half a code below H.273's 2^(depth-1), so it sees every chroma
   sample half a code high: |error| <= (0.5 rounding + 0.5 offset) x the channel gain 2(1-kb) ~ 1.9, plus the roundings
   of Y and of the output -- within 3 codes, and the MEAN error of B and R shows the bias (that is the decoder's doing;
   any other AVIF decoder uses 2^(depth-1)).
Also checks the down-filter definition on flat blocks."""
import numpy as np
import pytest

import cases
from avifgpu import abi


def round_trip_error(encoder, decoder, matrix_nclx, depth, host_depth, rng):
    """encode uint host RGB (already at code precision) to YCbCr 4:4:4 and decode back; returns max |delta| in codes."""
    w, h = 64, 32
    top = (1 << depth) - 1
    if depth == 8:
        rgb = rng.integers(0, 256, (h, w * 3)).astype(np.uint8)
        enc = abi.EncodeDesc(w, h, 8, 3, abi.ALPHA_NONE, 8, layout=abi.LAYOUT_PLANAR_YCBCR, chroma=abi.CHROMA_444, nclx=matrix_nclx)
        planes = encoder(enc, rgb)
        dec = abi.DecodeDesc(w, h, abi.COLORSPACE_YCBCR, abi.CHROMA_444, 8, abi.ALPHA_NONE, 8, matrix_nclx)
        back = decoder(dec, planes)
        return int(np.abs(back.astype(np.int32) - rgb.astype(np.int32)).max())
    # deeper images: feed codes through a float host with the clip transfer so the codes are exactly known
    codes = rng.integers(0, top + 1, (h, w * 3))
    rows = (codes.astype(np.float64) / top + 0.25 / top).astype(np.float32)  # trunc(v * max) == code
    enc = abi.EncodeDesc(w, h, 32, 3, abi.ALPHA_NONE, depth, abi.TRANSFER_CLIP, 80, abi.LAYOUT_PLANAR_YCBCR, abi.CHROMA_444, nclx=matrix_nclx)
    planes = encoder(enc, rows)
    dec = abi.DecodeDesc(w, h, abi.COLORSPACE_YCBCR, abi.CHROMA_444, depth, abi.ALPHA_NONE, 16, matrix_nclx)
    back = decoder(dec, planes).astype(np.float64) * top / 32768.0  # host 0..32768 -> code units
    return float(np.abs(back - codes).max())


MATRICES = [("601-default", None), ("709", cases.NCLX_709()), ("2020", cases.NCLX_2020_PQ()), ("derived-2020", cases.NCLX_DERIVED())]


@pytest.mark.parametrize("name,nclx", MATRICES, ids=[m[0] for m in MATRICES])
@pytest.mark.parametrize("depth", [8, 10, 12])
def test_round_trip_through_reference_decoder(port, ref, name, nclx, depth):
    err = round_trip_error(lambda d, r: port.encode(d, r), lambda d, p: ref.decode(d, p), nclx, depth, 8 if depth == 8 else 32,
                           np.random.default_rng(depth))
    assert err <= 3.0, err


def h273_forward(rgb, kr, kb, depth):
    """H.273 section 8.3 full range, float64: rgb = (..., 3) integer codes -> integer Y, Cb, Cr codes."""
    top = (1 << depth) - 1
    kg = 1.0 - kr - kb
    r, g, b = (rgb[..., i].astype(np.float64) for i in range(3))
    y = kr * r + kg * g + kb * b
    cb = (b - y) / (2 * (1 - kb))
    cr = (r - y) / (2 * (1 - kr))
    half = 1 << (depth - 1)
    q = lambda v: np.clip(np.floor(v + 0.5), 0, top).astype(np.int64)
    return q(y), q(cb + half), q(cr + half)


@pytest.mark.parametrize("name,nclx", MATRICES, ids=[m[0] for m in MATRICES])
@pytest.mark.parametrize("depth", [8, 10, 12])
def test_forward_matrix_is_h273(port, name, nclx, depth):
    rng = np.random.default_rng(depth * 7)
    w, h = 96, 64
    top = (1 << depth) - 1
    codes = rng.integers(0, top + 1, (h, w, 3))
    codes[0, :8] = [[top, 0, 0], [0, top, 0], [0, 0, top], [top, top, top], [0, 0, 0], [top, top, 0], [0, top, top], [top, 0, top]]
    if depth == 8:
        enc = abi.EncodeDesc(w, h, 8, 3, abi.ALPHA_NONE, 8, layout=abi.LAYOUT_PLANAR_YCBCR, chroma=abi.CHROMA_444, nclx=nclx)
        y, cb, cr, _ = port.encode(enc, codes.reshape(h, w * 3).astype(np.uint8))
    else:
        rows = (codes.reshape(h, w * 3).astype(np.float64) / top + 0.25 / top).astype(np.float32)
        enc = abi.EncodeDesc(w, h, 32, 3, abi.ALPHA_NONE, depth, abi.TRANSFER_CLIP, 80, abi.LAYOUT_PLANAR_YCBCR, abi.CHROMA_444, nclx=nclx)
        y, cb, cr, _ = port.encode(enc, rows)
    k = port.yuv_coefficients(nclx)
    ey, ecb, ecr = h273_forward(codes, float(k[0]), float(k[2]), depth)
    for got, expected in ((y, ey), (cb, ecb), (cr, ecr)):
        delta = np.abs(got.astype(np.int64) - expected)
        assert delta.max() <= 1                 # float32 vs float64 may straddle a .5 boundary ...
        assert (delta != 0).mean() < 2e-3       # ... on a handful of samples, never systematically
    # neutral grey sits on 2^(depth-1); saturated red / blue clip at the top code
    assert cb[0, 3] == cr[0, 3] == 1 << (depth - 1) and cb[0, 4] == 1 << (depth - 1)
    assert cr[0, 0] == top and cb[0, 2] == top
    # H.273 inverse (float64) of our codes returns the RGB codes within the quantisation bound
    kr, kb = float(k[0]), float(k[2])
    kg = 1 - kr - kb
    half = 1 << (depth - 1)
    yf, cbf, crf = y.astype(np.float64), cb.astype(np.float64) - half, cr.astype(np.float64) - half
    r = yf + 2 * (1 - kr) * crf
    b = yf + 2 * (1 - kb) * cbf
    g = (yf - kr * r - kb * b) / kg
    back = np.stack([r, g, b], axis=-1)
    inner = codes[1:]  # the constructed first row holds the clipped extremes
    assert np.abs(back[1:] - inner).max() <= 0.5 + 0.5 * 2 * (1 - min(kr, kb)) + 0.05


def test_identity_matrix_round_trip_is_lossless(port, ref):
    nclx = cases.NCLX_GBR()
    rng = np.random.default_rng(3)
    rgb = rng.integers(0, 256, (16, 48)).astype(np.uint8)
    enc = abi.EncodeDesc(16, 16, 8, 3, abi.ALPHA_NONE, 8, layout=abi.LAYOUT_PLANAR_YCBCR, chroma=abi.CHROMA_444, nclx=nclx)
    y, cb, cr, _ = port.encode(enc, rgb)
    px = rgb.reshape(16, 16, 3)
    assert np.array_equal(y, px[..., 1]) and np.array_equal(cb, px[..., 2]) and np.array_equal(cr, px[..., 0])  # GBR


def test_down_filter_definition_on_constructed_blocks(port):
    """BOX = mean of the float chroma of the block, TOP_LEFT = co-sited sample; odd edges use the samples present."""
    w, h = 5, 3
    rows = np.zeros((h, w * 3), np.uint8)
    rows[0, 0:3] = (255, 0, 0)   # only the top-left pixel of block (0,0) is red
    rows[2, 12:15] = (0, 0, 255)  # bottom-right corner pixel (a 1x1 block at the odd edge) is blue
    for down in (abi.DOWN_FILTER_BOX, abi.DOWN_FILTER_TOP_LEFT):
        enc = abi.EncodeDesc(w, h, 8, 3, abi.ALPHA_NONE, 8, layout=abi.LAYOUT_PLANAR_YCBCR, chroma=abi.CHROMA_420, down_filter=down)
        y, cb, cr, _ = port.encode(enc, rows)
        full = abi.EncodeDesc(w, h, 8, 3, abi.ALPHA_NONE, 8, layout=abi.LAYOUT_PLANAR_YCBCR, chroma=abi.CHROMA_444)
        _, cb444, cr444, _ = port.encode(full, rows)
        assert cb.shape == (2, 3)
        if down == abi.DOWN_FILTER_TOP_LEFT:
            assert cr[0, 0] == cr444[0, 0] and cb[1, 2] == cb444[2, 4]
        else:
            assert 128 < cr[0, 0] < cr444[0, 0]          # one red sample in four pulls Cr a quarter of the way
            assert cb[1, 2] == cb444[2, 4]               # a lone edge sample is its own mean


@pytest.mark.gpu
@pytest.mark.parametrize("depth", [8, 10, 12])
def test_gpu_round_trip_through_reference_decoder(gpu, checker, depth):
    err = round_trip_error(lambda d, r: gpu.encode(d, r), lambda d, p: checker.decode(d, p), cases.NCLX_2020_PQ(), depth,
                           8 if depth == 8 else 32, np.random.default_rng(depth + 10))
    assert err <= 3.0, err
