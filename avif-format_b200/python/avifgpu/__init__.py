"""Python binding (ctypes) of the avifgpu C ABI -- see include/avifgpu.h.  Filled in by runtime.py."""
from . import abi  # noqa: F401
