"""ctypes helpers for the packed kernel structs (cf. ``xopto/cl/cltypes.py``).

The reference mirrors every OpenCL scalar/vector type; the hot path only needs
natural-alignment ``Structure`` plus the scalar aliases used in plugin structs.
"""
import ctypes

Structure = ctypes.Structure
Array = ctypes.Array

cl_float = ctypes.c_float
cl_double = ctypes.c_double
cl_int = cl_int32_t = ctypes.c_int32
cl_uint = cl_uint32_t = ctypes.c_uint32
cl_long = cl_int64_t = ctypes.c_int64
cl_ulong = cl_uint64_t = ctypes.c_uint64


def raw_bytes(struct) -> bytes:
    """Bytes of a packed struct / struct array exactly as the kernel sees them."""
    return bytes(memoryview(struct).cast('B'))
