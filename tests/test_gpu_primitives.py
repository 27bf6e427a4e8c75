"""Primitive-level parity on a B200 (SURVEY.md component 4): every transfer function of ColorTransfer.cpp, the two HLG
OOTFs and the three libm functions underneath, GPU against the CPU checker, bit for bit, over dense sweeps of their
domains plus the IEEE special values.

One documented difference: WHICH NaN.  Where the result is a NaN, x86 delivers the "default NaN" 0xffc00000 (or the
quieted operand) and the GPU its canonical 0x7fffffff; the two are required to be NaN together, their payloads are
not compared.  No NaN reaches an output of the pixel path: integer outputs define NaN -> 0 (pixel_math.cuh FloatToCode)
and the decoders' float outputs are computed from integer codes."""
import numpy as np
import pytest

import cases
from avifgpu import abi

pytestmark = pytest.mark.gpu


def same_bits_or_both_nan(a, b):
    a = np.asarray(a, np.float32).ravel()
    b = np.asarray(b, np.float32).ravel()
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    return np.array_equal(nan_a, nan_b) and np.array_equal(a.view(np.uint32)[~nan_a], b.view(np.uint32)[~nan_a])


def sweep(lo, hi, count, seed):
    rng = np.random.default_rng(seed)
    uniform = rng.uniform(lo, hi, count).astype(np.float32)
    logarithmic = np.exp(rng.uniform(np.log(1e-12), np.log(max(hi, 1e-6)), count)).astype(np.float32)
    specials = np.array([0.0, -0.0, 1.0, -1.0, 0.5, np.inf, -np.inf, np.nan, 1e-45, 1e-38, 3.4e38, 1.0 / 12.0, 0.25, 125.0], np.float32)
    return np.concatenate([uniform, logarithmic, -logarithmic[: count // 8], specials])


@pytest.mark.parametrize("function,param,lo,hi", [
    (abi.FN_LINEAR_TO_PQ, 80.0, -0.5, 130.0), (abi.FN_LINEAR_TO_PQ, 10000.0, -0.5, 2.0), (abi.FN_PQ_TO_LINEAR, 80.0, -0.2, 1.2),
    (abi.FN_PQ_TO_LINEAR, 1000.0, -0.2, 1.2), (abi.FN_LINEAR_TO_SMPTE428, 0.0, -0.5, 2.0), (abi.FN_SMPTE428_TO_LINEAR, 0.0, -0.2, 1.2),
    (abi.FN_HLG_TO_LINEAR, 0.0, -0.2, 1.2), (abi.FN_LINEAR_TO_HLG, 0.0, -0.2, 1.2), (abi.FN_POWF, 0.2, 0.0, 4.0), (abi.FN_POWF, 2.6, 0.0, 4.0),
    (abi.FN_POWF, -1.5, 0.0, 4.0), (abi.FN_EXPF, 0.0, -100.0, 100.0), (abi.FN_LOGF, 0.0, 0.0, 1000.0)])
def test_transfer_functions(gpu, checker, function, param, lo, hi):
    x = sweep(lo, hi, 200000, function * 100 + int(param))
    expected = checker.transfer(function, x, param)
    got = gpu.transfer(function, x, param)
    assert same_bits_or_both_nan(expected, got), int((expected.view(np.uint32) != got.view(np.uint32)).sum())


@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("primaries", [abi.PRIMARIES_BT709, abi.PRIMARIES_BT2020])
def test_hlg_ootf(gpu, checker, inverse, primaries):
    rng = np.random.default_rng(17 + primaries)
    rgb = np.concatenate([rng.uniform(0.0, 1.0, (100000, 3)), np.exp(rng.uniform(np.log(1e-9), np.log(1000.0), (50000, 3))),
                          [[0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [0.0, 1.0, 0.0], [np.inf, 0.1, 0.2], [np.nan, 0.3, 0.3], [-0.1, 0.2, 0.3]]]).astype(np.float32)
    for gamma, peak in ((1.2, 1000.0), (1.0, 400.0), (1.5, 4000.0)):
        if inverse:
            expected = checker.hlg_inverse_ootf(rgb, primaries, gamma, peak)
        else:
            expected = checker.hlg_ootf(rgb, primaries, gamma, peak)
        got = gpu.hlg_ootf(rgb, primaries, gamma, peak, inverse=inverse)
        assert same_bits_or_both_nan(expected, got), (gamma, peak)
