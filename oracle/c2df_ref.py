"""Independent restatement of the reference's embedding codec and ``.c2df`` reader, used to
check the product's C++ batch parser and K1 kernel.  TEST INFRASTRUCTURE ONLY.

Follows, line by line:
  * src/filemaker.py:137-173  ``unpack_c2df``  (TLV walk)  +  :102-135 ``_load_entry``
  * src/search.py:16-22       ``l2n`` / ``dequantize_clip_u8``
  * src/search.py:24-41       ``decode_clip_from_c2df`` (error order and classes)
  * src/compress.py:76-86     ``quantize_u8_and_compress``
Pinned against the reference's own ``filemaker.py`` (importable in the build container) through
``tests/golden/make_golden.py`` and against the shipped ``apple.*`` fixtures (KAT-1..6).
zstd goes through the system libzstd with ctypes (``zstandard`` is not installed).
"""
from __future__ import annotations

import ctypes as C
import json
import struct

import numpy as np

_z = C.CDLL("libzstd.so.1")
_z.ZSTD_getFrameContentSize.restype = C.c_ulonglong
_z.ZSTD_getFrameContentSize.argtypes = [C.c_char_p, C.c_size_t]
_z.ZSTD_decompress.restype = C.c_size_t
_z.ZSTD_decompress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
_z.ZSTD_compress.restype = C.c_size_t
_z.ZSTD_compress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int]
_z.ZSTD_compressBound.restype = C.c_size_t
_z.ZSTD_compressBound.argtypes = [C.c_size_t]
_z.ZSTD_isError.restype = C.c_uint
_z.ZSTD_isError.argtypes = [C.c_size_t]


class OracleZstdError(Exception):
    pass


def zstd_decompress(buf: bytes) -> bytes:
    n = _z.ZSTD_getFrameContentSize(buf, len(buf))
    if n >= 2 ** 64 - 2:
        raise OracleZstdError("frame content size unknown / invalid frame")
    out = C.create_string_buffer(max(int(n), 1))
    r = _z.ZSTD_decompress(out, n, buf, len(buf))
    if _z.ZSTD_isError(r):
        raise OracleZstdError("zstd decode failed")
    return out.raw[:r]


def zstd_compress(buf: bytes, level: int = 19) -> bytes:
    cap = _z.ZSTD_compressBound(len(buf))
    out = C.create_string_buffer(cap)
    r = _z.ZSTD_compress(out, cap, buf, len(buf), level)
    assert not _z.ZSTD_isError(r)
    return out.raw[:r]


def _load_entry(t: int, payload: bytes):
    """filemaker.py:102-135: the value an entry stands for.  The reference builds it for EVERY entry of a file while it
    walks, so an entry that does not load (bad UTF-8, bad JSON, an array whose bytes do not fit its shape, an
    unknown type code, a payload shorter than its fixed size) makes the whole file unreadable."""
    if t == 6:
        return None
    if t == 7:
        return bool(struct.unpack_from("<B", payload, 0)[0])
    if t == 2:
        return struct.unpack_from("<q", payload, 0)[0]
    if t == 3:
        return struct.unpack_from("<d", payload, 0)[0]
    if t in (0, 1, 4):
        (n,) = struct.unpack_from("<I", payload, 0)
        inner = payload[4:4 + n]
        return inner if t == 0 else inner.decode("utf-8") if t == 1 else json.loads(inner.decode("utf-8"))
    if t == 5:
        off = 0
        (dt_len,) = struct.unpack_from("<B", payload, off)
        off += 1
        dt = payload[off:off + dt_len].decode("utf-8")
        off += dt_len
        (ndim,) = struct.unpack_from("<B", payload, off)
        off += 1
        shape = []
        for _ in range(ndim):
            (d,) = struct.unpack_from("<I", payload, off)
            off += 4
            shape.append(int(d))
        (data_len,) = struct.unpack_from("<I", payload, off)
        off += 4
        return np.frombuffer(payload[off:off + data_len], dtype=np.dtype(dt)).reshape(shape)
    raise ValueError(f"unknown type code: {t}")


def walk(data: bytes):
    """Entries of a .c2df as {key: (type_code, raw_payload)} + header dict (filemaker.py:137-173).  Every entry is
    loaded on the way, as the reference does (the loaded values are dropped: callers here want the raw payloads)."""
    if data[:4] != b"C2DF":
        raise AssertionError("bad magic")
    off = 4
    struct.unpack_from("<H", data, off)        # version: read and ignored (filemaker.py:147)
    off += 2
    (hlen,) = struct.unpack_from("<I", data, off)
    off += 4
    header = json.loads(data[off:off + hlen].decode("utf-8")) if hlen > 0 else {}
    off += hlen
    (n_items,) = struct.unpack_from("<I", data, off)
    off += 4
    out = {}
    for _ in range(n_items):
        (klen,) = struct.unpack_from("<H", data, off)
        off += 2
        key = data[off:off + klen].decode("utf-8")
        off += klen
        (t,) = struct.unpack_from("<B", data, off)
        off += 1
        if t in (2, 3):
            size = 8
        elif t == 7:
            size = 1
        elif t == 6:
            size = 0
        else:
            (size,) = struct.unpack_from("<I", data, off)
            off += 4
        payload = data[off:off + size]         # slices clamp at the end of the data; unpack_from raises
        off += size
        _load_entry(t, payload)
        out[key] = (t, payload)
    return out, header


def l2n(x, eps=1e-9):
    n = np.linalg.norm(x, axis=-1, keepdims=True)
    return x / np.maximum(n, eps)


def dequantize_clip_u8(q):
    z = (q.astype(np.float32) / 255.0) * 2.0 - 1.0
    return l2n(z.astype(np.float32))


def quantize_u8(z_unit):
    return np.clip(np.round((z_unit * 0.5 + 0.5) * 255.0), 0, 255).astype(np.uint8)


def decode_clip(data: bytes):
    """bytes of one .c2df → (u8 codes, fp32 unit vector).  Raises what the reference raises."""
    entries, _ = walk(data)
    if "clip_stream" not in entries or "clip_meta" not in entries:
        raise ValueError("No 'clip_stream' or 'clip_meta'")
    t_meta, p_meta = entries["clip_meta"]
    if t_meta == 4:
        (n,) = struct.unpack_from("<I", p_meta, 0)
        meta = json.loads(p_meta[4:4 + n].decode("utf-8")) or {}
    elif t_meta == 6:
        meta = {}
    else:
        raise AttributeError("clip_meta has no .get")
    dim = int(meta.get("dim", 0))
    if dim <= 0:
        raise ValueError("Invalid clip_meta.dim")
    t_s, p_s = entries["clip_stream"]
    if t_s != 0:
        raise TypeError("clip_stream is not bytes")
    (n,) = struct.unpack_from("<I", p_s, 0)
    q = np.frombuffer(zstd_decompress(p_s[4:4 + n]), dtype=np.uint8)
    if q.size != dim:
        raise ValueError("Dimension didn't match")
    return q, dequantize_clip_u8(q).astype(np.float32)


# ---- IxFI --------------------------------------------------------------
def read_ixfi(path):
    raw = open(path, "rb").read()
    assert raw[:4] == b"IxFI"
    d, ntotal, _d0, _d1, trained, metric, count = struct.unpack_from("<iqqqBiQ", raw, 4)
    assert count == ntotal * d and metric == 0 and trained == 1
    x = np.frombuffer(raw, dtype="<f4", count=count, offset=45).reshape(ntotal, d)
    return x


def write_ixfi(path, x):
    x = np.ascontiguousarray(x, dtype="<f4")
    n, d = x.shape
    with open(path, "wb") as f:
        f.write(b"IxFI" + struct.pack("<iqqqBiQ", d, n, 1 << 20, 1 << 20, 1, 0, n * d))
        f.write(x.tobytes())
