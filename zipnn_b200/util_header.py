"""Header enums of the ZipNN stream (mirrors reference zipnn/util_header.py:5-44).

Values are part of the on-disk format: header[7] = method, header[8] = input
format, header[10] = lossy type.  Strings are accepted case-insensitively, as in
the reference (`EnumMethod("huffman")`).
"""
from enum import Enum


class _CaseInsensitive(Enum):
    @classmethod
    def _missing_(cls, value):
        if isinstance(value, str):
            return cls.__members__.get(value.upper())
        return None


class EnumMethod(_CaseInsensitive):
    AUTO = 0
    HUFFMAN = 1
    ZSTD = 2
    LZ4 = 3
    SNAPPY = 4


class EnumFormat(_CaseInsensitive):
    BYTE = 1
    TORCH = 2
    NUMPY = 3
    FILE = 4


class EnumLossy(_CaseInsensitive):
    NONE = 0
    INTEGER = 1
    UNSIGN = 2
