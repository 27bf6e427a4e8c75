"""The colour-profile step for matrix profiles: the ICC parser + matrix construction of
avifgpu_icc_to_rec2020_linear_matrix against an independent float64 construction, and its refusals.  The profiles are built
here from the ICC.1 tag layouts (header, tag table, 'XYZ ' and 'curv' / 'para' tags); lcms2 is not in the reference tree,
so the stage it replaces is unpinned -- what IS pinned is the colour science: BT.2087's Rec.709 -> Rec.2020 matrix."""
import struct

import numpy as np
import pytest

import avifgpu
from avifgpu import abi

D50 = np.array([0.9642, 1.0, 0.8249])
BRADFORD = np.array([[0.8951, 0.2664, -0.1614], [-0.7502, 1.7135, 0.0367], [0.0389, -0.0685, 1.0296]])


def rgb_to_xyz(primaries, white):
    xyz = np.array([[x / y, 1.0, (1 - x - y) / y] for x, y in primaries]).T
    w = np.array([white[0] / white[1], 1.0, (1 - white[0] - white[1]) / white[1]])
    return xyz * np.linalg.solve(xyz, w)


def adapt(white_xy, target=D50):
    w = np.array([white_xy[0] / white_xy[1], 1.0, (1 - white_xy[0] - white_xy[1]) / white_xy[1]])
    return np.linalg.inv(BRADFORD) @ np.diag((BRADFORD @ target) / (BRADFORD @ w)) @ BRADFORD


def s15f16(v):
    return struct.pack(">i", int(round(v * 65536.0)))


def make_profile(primaries, white, curve="curv0", lut=False, colour_space=b"RGB "):
    colorants = adapt(white) @ rgb_to_xyz(primaries, white)  # PCS-adapted, as a v4 profile stores them
    tags = []
    for c, name in enumerate((b"rXYZ", b"gXYZ", b"bXYZ")):
        tags.append((name, b"XYZ \0\0\0\0" + b"".join(s15f16(colorants[r, c]) for r in range(3))))
    if curve == "curv0":
        trc = b"curv\0\0\0\0" + struct.pack(">I", 0)
    elif curve == "gamma1":
        trc = b"curv\0\0\0\0" + struct.pack(">I", 1) + struct.pack(">H", 0x0100) + b"\0\0"
    elif curve == "gamma2.2":
        trc = b"curv\0\0\0\0" + struct.pack(">I", 1) + struct.pack(">H", int(2.2 * 256)) + b"\0\0"
    else:
        trc = b"para\0\0\0\0" + struct.pack(">HH", 0, 0) + s15f16(1.0)
    for name in (b"rTRC", b"gTRC", b"bTRC"):
        tags.append((name, trc))
    if lut:
        tags.append((b"A2B0", b"mft2" + b"\0" * 60))
    header = bytearray(128)
    header[12:16] = b"mntr"
    header[16:20] = colour_space
    header[20:24] = b"XYZ "
    header[36:40] = b"acsp"
    table = struct.pack(">I", len(tags))
    offset = 128 + 4 + 12 * len(tags)
    body = b""
    for name, data in tags:
        data = data + b"\0" * (-len(data) % 4)
        table += name + struct.pack(">II", offset + len(body), len(data))
        body += data
    blob = bytes(header) + table + body
    return blob[:0] + struct.pack(">I", len(blob)) + blob[4:]


REC709 = [(0.64, 0.33), (0.30, 0.60), (0.15, 0.06)]
REC2020 = [(0.708, 0.292), (0.170, 0.797), (0.131, 0.046)]
P3 = [(0.680, 0.320), (0.265, 0.690), (0.150, 0.060)]
D65 = (0.3127, 0.3290)


@pytest.mark.parametrize("curve", ["curv0", "gamma1", "para"])
def test_linear_rec709_profile_gives_the_bt2087_matrix(curve):
    matrix, same = avifgpu.icc_to_rec2020_linear_matrix(make_profile(REC709, D65, curve))
    assert not same
    bt2087 = np.array([[0.6274, 0.3293, 0.0433], [0.0691, 0.9195, 0.0114], [0.0164, 0.0880, 0.8956]])
    assert np.abs(matrix - bt2087).max() < 2e-4
    # and the independent float64 construction, to the precision the s15Fixed16 colorants allow
    expected = np.linalg.inv(adapt(D65) @ rgb_to_xyz(REC2020, D65)) @ (adapt(D65) @ rgb_to_xyz(REC709, D65))
    assert np.abs(matrix - expected).max() < 5e-5
    assert np.abs(matrix.sum(axis=1) - 1.0).max() < 5e-5  # white maps to white


def test_rec2020_profile_is_recognised_and_p3_is_not():
    matrix, same = avifgpu.icc_to_rec2020_linear_matrix(make_profile(REC2020, D65))
    assert same and np.abs(matrix - np.eye(3)).max() < 2e-4
    matrix, same = avifgpu.icc_to_rec2020_linear_matrix(make_profile(P3, D65))
    assert not same and matrix[0, 0] < 0.8


def test_profiles_the_matrix_cannot_express_are_refused():
    for blob in (make_profile(REC709, D65, "gamma2.2"), make_profile(REC709, D65, lut=True), make_profile(REC709, D65, colour_space=b"CMYK")):
        with pytest.raises(avifgpu.AvifGpuError) as info:
            avifgpu.icc_to_rec2020_linear_matrix(blob)
        assert info.value.status == abi.ERR_UNSUPPORTED
    with pytest.raises(avifgpu.AvifGpuError) as info:
        avifgpu.icc_to_rec2020_linear_matrix(b"not a profile" * 20)
    assert info.value.status == abi.ERR_BAD_PARAM


def test_row_matrix_is_applied_before_everything_else(port):
    """Restatement: matrix first (alpha untouched), then the reference's clamp / premultiply / curve -- an identity matrix
    changes nothing, a channel swap swaps the codes of the interleaved layout."""
    import cases
    w, h = 24, 6
    rows = cases.float_host_rows(np.random.default_rng(8), h, w, 4)
    desc = abi.EncodeDesc(w, h, 32, 4, abi.ALPHA_STRAIGHT, 12, abi.TRANSFER_PQ, 80)
    plain = port.encode(desc, rows)[0]
    desc.row_matrix_enabled = 1
    desc.row_matrix = type(desc.row_matrix)(1, 0, 0, 0, 1, 0, 0, 0, 1)
    assert np.array_equal(port.encode(desc, rows)[0], plain)
    desc.row_matrix = type(desc.row_matrix)(0, 0, 1, 0, 1, 0, 1, 0, 0)
    swapped = port.encode(desc, rows)[0].reshape(h, w, 4)
    assert np.array_equal(swapped[..., [2, 1, 0, 3]], plain.reshape(h, w, 4))
