"""dtype table and shape packing of the ZipNN stream header.

Mirrors reference zipnn/util_torch.py:89-159 (shape packing) and :176-234
(`ZipNNDtypeEnum`; the integer code is header byte 15).
"""
import struct

import numpy as np
import torch

from .util_header import EnumFormat

# code -> (name, torch dtype, numpy dtype); codes are the reference's (util_torch.py:177-207)
_DTYPES = [
    (0, "none", None, None),
    (1, "float32", torch.float32, np.float32),
    (2, "float", torch.float32, np.float32),
    (3, "float64", torch.float64, np.float64),
    (4, "float16", torch.float16, np.float16),
    (5, "half", torch.float16, np.float16),
    (6, "bfloat16", torch.bfloat16, None),
    (13, "uint8", torch.uint8, np.uint8),
    (14, "uint16", None, np.uint16),
    (15, "uint32", None, np.uint32),
    (16, "uint64", None, np.uint64),
    (17, "int8", torch.int8, np.int8),
    (18, "int16", torch.int16, np.int16),
    (20, "int32", torch.int32, np.int32),
    (22, "int64", torch.int64, np.int64),
    (24, "bool", torch.bool, np.bool_),
    (29, "float8_e4m3fn", torch.float8_e4m3fn, None),
    (30, "float8_e5m2", torch.float8_e5m2, None),
]

FLOAT32, FLOAT, FLOAT64, FLOAT16, HALF, BFLOAT16 = 1, 2, 3, 4, 5, 6
UINT32 = 15
FLOAT8_E4M3FN, FLOAT8_E5M2 = 29, 30


def dtype_code(dtype) -> int:
    """Header code for a torch dtype, numpy dtype or dtype string (first match wins,
    as in `ZipNNDtypeEnum.from_dtype`: torch.float32 -> 1, "float" -> 2, "half" -> 5)."""
    if isinstance(dtype, str):
        key = dtype.lower()
        for code, name, _, _ in _DTYPES:
            if key == name:
                return code
        return 0
    for code, _, tdt, ndt in _DTYPES:
        if tdt is not None and dtype == tdt:
            return code
    for code, _, tdt, ndt in _DTYPES:
        try:
            if ndt is not None and np.dtype(dtype) == np.dtype(ndt):
                return code
        except TypeError:
            pass
    return 0


def dtype_name(code: int) -> str:
    for c, name, _, _ in _DTYPES:
        if c == code:
            return name
    return "none"


def torch_dtype_of_code(code: int):
    for c, _, tdt, _ in _DTYPES:
        if c == code:
            return tdt
    return None


def zipnn_pack_shape(shape) -> bytes:
    """[ndim] then per dim [k in {1,2,4,8}][dim as k little-endian bytes] (util_torch.py:89-118)."""
    out = bytearray([len(shape)])
    for dim in shape:
        dim = int(dim)
        if dim < 1 << 8:
            out += b"\x01" + struct.pack("<B", dim)
        elif dim < 1 << 16:
            out += b"\x02" + struct.pack("<H", dim)
        elif dim < 1 << 32:
            out += b"\x04" + struct.pack("<I", dim)
        else:
            out += b"\x08" + struct.pack("<Q", dim)
    return bytes(out)


def zipnn_unpack_shape(packed):
    """-> (shape tuple, bytes consumed) (util_torch.py:121-159).  The reference
    forgets to count the 8 payload bytes of an 8-byte dimension; we count them, which
    is what its writer means (dims >= 2**32 do not occur in practice)."""
    packed = bytes(packed[:1 + 9 * 255])
    ndim = packed[0]
    dims, i = [], 1
    while len(dims) < ndim:
        k = packed[i]
        i += 1
        if k not in (1, 2, 4, 8):
            raise ValueError("corrupt shape descriptor in ZipNN header")
        dims.append(int.from_bytes(packed[i:i + k], "little"))
        i += k
    return tuple(dims), i


def zipnn_is_floating_point(data_format_value, data, bytearray_dtype) -> bool:
    """util_torch.py:162-168"""
    if data_format_value == EnumFormat.TORCH.value:
        return torch.is_floating_point(data)
    if data_format_value == EnumFormat.NUMPY.value:
        return np.issubdtype(data.dtype, np.floating)
    if data_format_value == EnumFormat.BYTE.value:
        name = bytearray_dtype if isinstance(bytearray_dtype, str) else dtype_name(dtype_code(bytearray_dtype))
        return name in ("float64", "float32", "float16", "bfloat16", "float8_e4m3fn", "float8_e5m2")
    return False
