"""Pins this project's definition of the libheif stage (forward matrix + quantisation) by a
round trip through the reference's own decoder: RGB codes -> (forward, 4:4:4) -> Y/Cb/Cr codes -> the compiled
reference's DecodeYUV*Row* must return the RGB codes within +-2 codes (the B channel's gain 2(1-kb) ~ 1.8 amplifies
the +-1/2-code rounding of Cb).  Also checks the down-filter definition on flat blocks."""
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
    assert err <= 2.0, err


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
    assert err <= 2.0, err
