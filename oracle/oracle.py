"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes loaders for the two CPU checkers:

  load_restatement()   oracle/liboracle.so       the plain-C restatement (oracle/avif_oracle.c)
  load_reference()     oracle/_ref/libavifref.so the UNMODIFIED reference translation units compiled in place
                                                 (oracle/Makefile), or None when it has not been built

Both expose the same Python surface (class CpuChecker) so a test can run either against the GPU library.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module;
the product package never does.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_PKG_PY = os.path.join(_ROOT, "avif-format_b200", "python")
if _PKG_PY not in sys.path:
    sys.path.insert(0, _PKG_PY)

from avifgpu import abi  # noqa: E402  (declarations only, loads no native code)


class OracleError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"status {status}: {message}")
        self.status = status


def build(quiet=True):
    """Compile liboracle.so and, when /root/reference is present, _ref/libavifref.so."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


class CpuChecker:
    """Uniform Python surface over liboracle.so (prefix avif_oracle_) or libavifref.so (prefix avifref_)."""

    def __init__(self, path, prefix, kind):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.kind = kind  # "port" (restatement) or "reference" (compiled reference TUs)
        self.path = path
        f = self._fn
        f("last_error", C.c_char_p, [])
        f("libm_version", C.c_char_p, [])
        f("transfer_f32", C.c_int, [C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_size_t])
        f("hlg_ootf", C.c_int, [C.c_void_p, C.c_size_t, C.c_int32, C.c_float, C.c_float])
        f("hlg_inverse_ootf", C.c_int, [C.c_void_p, C.c_size_t, C.c_int32, C.c_float, C.c_float])
        f("premultiply_u8", C.c_uint8, [C.c_uint8, C.c_uint8])
        f("premultiply_u16", C.c_uint16, [C.c_uint16, C.c_uint16, C.c_uint16])
        f("premultiply_f32", C.c_float, [C.c_float, C.c_float, C.c_float])
        f("unpremultiply_u8", C.c_uint8, [C.c_uint8, C.c_uint8])
        f("unpremultiply_u16", C.c_uint16, [C.c_uint16, C.c_uint16, C.c_uint16])
        f("unpremultiply_f32", C.c_float, [C.c_float, C.c_float, C.c_float])
        f("premultiply_table_u16", None, [C.c_uint16, C.c_int, C.c_void_p])
        f("premultiply_table_u8", None, [C.c_int, C.c_void_p])
        f("get_yuv_coefficients", C.c_int, [C.POINTER(abi.Nclx), C.c_void_p])
        f("get_hlg_luma_coefficients", C.c_int, [C.c_int32, C.c_void_p])
        f("build_yuv_tables", C.c_int, [C.POINTER(abi.Nclx), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p])
        f("encode_image", C.c_int, [C.POINTER(abi.EncodeDesc), C.c_void_p, C.c_int64, C.POINTER(abi.Planes)])
        f("decode_image", C.c_int, [C.POINTER(abi.DecodeDesc), C.POINTER(abi.Planes), C.c_void_p, C.c_int64])
        f("encode_image_mt", C.c_int, [C.POINTER(abi.EncodeDesc), C.c_void_p, C.c_int64, C.POINTER(abi.Planes), C.c_int32])
        f("decode_image_mt", C.c_int, [C.POINTER(abi.DecodeDesc), C.POINTER(abi.Planes), C.c_void_p, C.c_int64, C.c_int32])
        if prefix == "avif_oracle_":
            f("build_depth_lut", C.c_int, [C.c_int32, C.c_int32, C.c_void_p])
            f("rgb_codes_to_ycbcr_mt", C.c_int, [C.POINTER(abi.EncodeDesc), C.c_void_p, C.c_int64, C.POINTER(abi.Planes), C.c_int32])

    def _fn(self, name, restype, argtypes):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype = restype
        fn.argtypes = argtypes
        setattr(self, "_" + name, fn)

    def _check(self, status):
        if status != 0:
            raise OracleError(status, self._last_error().decode("utf-8", "replace"))

    # ---- scalars ----------------------------------------------------------------------------------------
    def libm_version(self):
        return self._libm_version().decode()

    def transfer(self, function, values, param=0.0):
        values = np.ascontiguousarray(values, dtype=np.float32)
        out = np.empty_like(values)
        self._check(self._transfer_f32(function, param, values.ctypes.data, out.ctypes.data, values.size))
        return out

    def hlg_ootf(self, rgb, primaries, gamma, peak):
        rgb = np.ascontiguousarray(rgb, dtype=np.float32).copy()
        self._check(self._hlg_ootf(rgb.ctypes.data, rgb.size // 3, primaries, gamma, peak))
        return rgb

    def hlg_inverse_ootf(self, rgb, primaries, gamma, peak):
        rgb = np.ascontiguousarray(rgb, dtype=np.float32).copy()
        self._check(self._hlg_inverse_ootf(rgb.ctypes.data, rgb.size // 3, primaries, gamma, peak))
        return rgb

    def premultiply_table(self, max_value, unpremultiply=False):
        if max_value == 255:
            out = np.empty((256, 256), np.uint8)
            self._premultiply_table_u8(int(unpremultiply), out.ctypes.data)
        else:
            out = np.empty((max_value + 1, max_value + 1), np.uint16)
            self._premultiply_table_u16(max_value, int(unpremultiply), out.ctypes.data)
        return out

    def yuv_coefficients(self, nclx):
        out = np.zeros(3, np.float32)
        self._check(self._get_yuv_coefficients(C.byref(nclx) if nclx is not None else None, out.ctypes.data))
        return out

    def hlg_luma_coefficients(self, primaries):
        out = np.zeros(3, np.float32)
        self._check(self._get_hlg_luma_coefficients(primaries, out.ctypes.data))
        return out

    def yuv_tables(self, nclx, bit_depth, monochrome, has_alpha=True):
        n = 1 << bit_depth
        y = np.zeros(n, np.float32)
        uv = None if monochrome else np.zeros(n, np.float32)
        a = np.zeros(n, np.float32) if has_alpha else None
        self._check(self._build_yuv_tables(C.byref(nclx) if nclx is not None else None, bit_depth, int(monochrome),
                                           y.ctypes.data, uv.ctypes.data if uv is not None else None,
                                           a.ctypes.data if a is not None else None))
        return y, uv, a

    def depth_lut(self, host_depth, image_bit_depth):
        out = np.zeros(256 if host_depth == 8 else 32769, np.uint16)
        self._check(self._build_depth_lut(host_depth, image_bit_depth, out.ctypes.data))
        return out

    # ---- images -----------------------------------------------------------------------------------------
    def encode(self, desc, rows, threads=1, pad=0, planes=None):
        """rows: (H, W*channels) host array.  Returns the list of 4 plane arrays (None where absent)."""
        rows = np.ascontiguousarray(rows)
        assert rows.dtype == abi.host_dtype(desc.host_depth)
        assert rows.shape == (desc.height, desc.width * desc.host_channels), rows.shape
        if planes is None:
            planes = alloc_planes(abi.encode_plane_shapes(desc), abi.code_dtype(desc.image_bit_depth), pad)
        p = abi.planes_from_arrays(planes)
        status = self._encode_image_mt(C.byref(desc), rows.ctypes.data, rows.strides[0] if rows.size else 0,
                                       C.byref(p), threads)
        self._check(status)
        return planes

    def decode(self, desc, planes, threads=1):
        """planes: list of 4 arrays/None (2-D).  Returns the (H, W*channels) host array."""
        channels = abi.decode_host_channels(desc)
        rows = np.zeros((desc.height, desc.width * channels), abi.host_dtype(desc.host_depth))
        p = abi.planes_from_arrays(planes)
        status = self._decode_image_mt(C.byref(desc), C.byref(p), rows.ctypes.data, rows.strides[0] if rows.size else 0,
                                       threads)
        self._check(status)
        return rows

    def rgb_codes_to_ycbcr(self, desc, interleaved, pad=0, threads=1, planes=None):
        interleaved = np.ascontiguousarray(interleaved)
        if planes is None:
            planes = alloc_planes(abi.encode_plane_shapes(desc), abi.code_dtype(desc.image_bit_depth), pad)
        p = abi.planes_from_arrays(planes)
        self._check(self._rgb_codes_to_ycbcr_mt(C.byref(desc), interleaved.ctypes.data, interleaved.strides[0], C.byref(p),
                                                threads))
        return planes


def alloc_planes(shapes, dtype, pad=0, fill=0xCD):
    """Planes with `pad` extra samples of row padding (views into wider buffers), filled with a sentinel."""
    out = []
    for shape in shapes:
        if shape is None:
            out.append(None)
            continue
        rows, cols = shape
        backing = np.full((max(rows, 0), cols + pad), fill, dtype=dtype)
        out.append(backing[:, :cols])
    return out


_cache = {}


def load_restatement():
    if "port" not in _cache:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _cache["port"] = CpuChecker(path, "avif_oracle_", "port")
    return _cache["port"]


def load_reference():
    """The compiled reference, or None if oracle/_ref/libavifref.so is absent and cannot be built here."""
    if "reference" not in _cache:
        path = os.path.join(_HERE, "_ref", "libavifref.so")
        if not os.path.exists(path) and os.path.exists("/root/reference/src/common/ColorTransfer.cpp"):
            build()
        _cache["reference"] = CpuChecker(path, "avifref_", "reference") if os.path.exists(path) else None
    return _cache["reference"]


def best_checker():
    """The compiled reference when available, else the restatement."""
    return load_reference() or load_restatement()
