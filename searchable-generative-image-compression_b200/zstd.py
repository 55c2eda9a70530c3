"""Minimal one-shot zstd (ctypes over the system ``libzstd.so.1``).

The reference uses the ``zstandard`` wheel (requirements.txt:30; src/compress.py:66,78 —
``ZstdCompressor(level=19).compress``; src/search.py:35 — ``ZstdDecompressor().decompress``).
That wheel is not in this image but the shared library it wraps is, so the two calls the
path needs are bound directly.  ``decompress`` keeps python-zstandard's contract: the frame
header must carry the content size, otherwise :class:`ZstdError`.
"""
from __future__ import annotations

import ctypes as C

_CONTENTSIZE_UNKNOWN = 2 ** 64 - 1
_CONTENTSIZE_ERROR = 2 ** 64 - 2


class ZstdError(Exception):
    pass


_z = None


def _lib():
    global _z
    if _z is None:
        last = None
        for name in ("libzstd.so.1", "libzstd.so"):
            try:
                z = C.CDLL(name)
                break
            except OSError as e:
                last = e
        else:
            raise ImportError(f"libzstd not found: {last}")
        z.ZSTD_compressBound.restype = C.c_size_t
        z.ZSTD_compressBound.argtypes = [C.c_size_t]
        z.ZSTD_compress.restype = C.c_size_t
        z.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        z.ZSTD_decompress.restype = C.c_size_t
        z.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        z.ZSTD_getFrameContentSize.restype = C.c_ulonglong
        z.ZSTD_getFrameContentSize.argtypes = [C.c_void_p, C.c_size_t]
        z.ZSTD_isError.restype = C.c_uint
        z.ZSTD_isError.argtypes = [C.c_size_t]
        z.ZSTD_getErrorName.restype = C.c_char_p
        z.ZSTD_getErrorName.argtypes = [C.c_size_t]
        _z = z
    return _z


def compress(data: bytes, level: int = 19) -> bytes:
    z = _lib()
    data = bytes(data)
    cap = z.ZSTD_compressBound(len(data))
    buf = C.create_string_buffer(cap)
    n = z.ZSTD_compress(buf, cap, data, len(data), int(level))
    if z.ZSTD_isError(n):
        raise ZstdError(z.ZSTD_getErrorName(n).decode())
    return buf.raw[:n]


def decompress(frame: bytes) -> bytes:
    z = _lib()
    if not isinstance(frame, (bytes, bytearray, memoryview)):
        raise TypeError("a bytes-like object is required")
    frame = bytes(frame)
    size = z.ZSTD_getFrameContentSize(frame, len(frame))
    if size == _CONTENTSIZE_ERROR:
        raise ZstdError("error determining content size from frame header")
    if size == _CONTENTSIZE_UNKNOWN:
        raise ZstdError("could not determine content size in frame header")
    if size == 0:
        return b""
    buf = C.create_string_buffer(size)
    n = z.ZSTD_decompress(buf, size, frame, len(frame))
    if z.ZSTD_isError(n):
        raise ZstdError("decompression error: " + z.ZSTD_getErrorName(n).decode())
    return buf.raw[:n]
