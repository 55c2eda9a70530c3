"""``.c2df`` container — host-side mirror of the reference's
``src/filemaker.py`` (``pack_c2df`` :75-100, ``unpack_c2df`` :137-173).

Little-endian TLV::

    "C2DF" | ver:u16 | hlen:u32 | header JSON | n_items:u32 |
    n_items x { klen:u16 | key utf-8 | type:u8 | payload }

Scalars are stored inline (INT ``<q``, FLOAT ``<d``, BOOL 1 byte, NONE nothing); every
other type is ``plen:u32`` + payload, and BYTES / STR / JSON payloads begin with their own
``u32`` length.  NP payload = ``dtlen:u8 | dtype.str | ndim:u8 | ndim x u32 | nbytes:u32 | data``.
There is no index table: readers walk the entries in order.

The batched, multi-threaded reader used for ingest is the C++ one in
``csrc/c2df_walk.cpp``; this module is the single-file path (``query-c2df``) and the
writer used to synthesise corpora.
"""
from __future__ import annotations

import json
import struct
import sys
from pathlib import Path

import numpy as np

BYTES, STR, INT, FLOAT, JSON, NP, NONE, BOOL = range(8)
_INLINE_SIZE = {INT: 8, FLOAT: 8, BOOL: 1, NONE: 0}
_INT_KEYS = frozenset({"token_length", "num_tokens", "n_tokens"})
MAGIC = b"C2DF"


def _np_payload(arr: np.ndarray) -> bytes:
    dt = arr.dtype.str.encode("utf-8")
    raw = arr.tobytes(order="C")
    dims = b"".join(struct.pack("<I", int(s)) for s in arr.shape)
    return struct.pack("<B", len(dt)) + dt + struct.pack("<B", arr.ndim) + dims + struct.pack("<I", len(raw)) + raw


def _lenpref(b: bytes) -> bytes:
    return struct.pack("<I", len(b)) + b


def _as_array(val):
    if isinstance(val, np.ndarray):
        return val
    torch = sys.modules.get("torch")
    if torch is not None and isinstance(val, torch.Tensor):
        return val.detach().cpu().contiguous().numpy()
    return None


def _encode(key: str, val):
    """(type code, payload) for one entry; same precedence as filemaker._dump_entry."""
    if key.endswith("_shape"):  # filemaker.py:22-32 — shape vectors are forced to int32 arrays
        return NP, _np_payload(np.asarray(val, dtype=np.int32))
    if key in _INT_KEYS or key.endswith("_length"):  # filemaker.py:35-36
        return INT, struct.pack("<q", int(val))
    if val is None:
        return NONE, b""
    if isinstance(val, bool):
        return BOOL, struct.pack("<B", int(val))
    if isinstance(val, int):
        return INT, struct.pack("<q", val)
    if isinstance(val, float):
        return FLOAT, struct.pack("<d", val)
    if isinstance(val, (bytes, bytearray, memoryview)):
        return BYTES, _lenpref(bytes(val))
    if isinstance(val, str):
        return STR, _lenpref(val.encode("utf-8"))
    arr = _as_array(val)
    if arr is not None:
        return NP, _np_payload(arr)
    if isinstance(val, (list, dict)):
        return JSON, _lenpref(json.dumps(val, ensure_ascii=False).encode("utf-8"))
    return STR, _lenpref(str(val).encode("utf-8"))


def pack_c2df(enc_result: dict, header: dict) -> bytes:
    out = [MAGIC, struct.pack("<H", int(header.get("version", 2)))]
    out.append(_lenpref(json.dumps(header, ensure_ascii=False).encode("utf-8")))
    out.append(struct.pack("<I", len(enc_result)))
    for key, val in enc_result.items():
        kb = key.encode("utf-8")
        code, payload = _encode(key, val)
        out.append(struct.pack("<H", len(kb)) + kb + struct.pack("<B", code))
        out.append(payload if code in _INLINE_SIZE else _lenpref(payload))
    return b"".join(out)


def _decode(code: int, payload: bytes):
    if code == NONE:
        return None
    if code == BOOL:
        return bool(struct.unpack_from("<B", payload)[0])
    if code == INT:
        return struct.unpack_from("<q", payload)[0]
    if code == FLOAT:
        return struct.unpack_from("<d", payload)[0]
    if code in (BYTES, STR, JSON):
        (n,) = struct.unpack_from("<I", payload)
        body = payload[4:4 + n]
        if code == BYTES:
            return body
        text = body.decode("utf-8")
        return text if code == STR else json.loads(text)
    if code == NP:
        pos = 1 + payload[0]
        dt = np.dtype(payload[1:pos].decode("utf-8"))
        ndim = payload[pos]
        pos += 1
        shape = struct.unpack_from(f"<{ndim}I", payload, pos)
        pos += 4 * ndim
        (nbytes,) = struct.unpack_from("<I", payload, pos)
        pos += 4
        return np.frombuffer(payload[pos:pos + nbytes], dtype=dt).reshape([int(s) for s in shape])
    raise ValueError(f"unknown type code: {code}")


def unpack_c2df(src):
    """``(enc_result, header)`` from a path or a bytes-like object."""
    data = Path(src).read_bytes() if isinstance(src, (str, Path)) else bytes(src)
    assert data[:4] == MAGIC, "bad magic"
    _ver, hlen = struct.unpack_from("<HI", data, 4)
    pos = 10
    header = json.loads(data[pos:pos + hlen].decode("utf-8")) if hlen > 0 else {}
    pos += hlen
    (n_items,) = struct.unpack_from("<I", data, pos)
    pos += 4
    entries = {}
    for _ in range(n_items):
        (klen,) = struct.unpack_from("<H", data, pos)
        pos += 2
        key = data[pos:pos + klen].decode("utf-8")
        pos += klen
        code = data[pos]
        pos += 1
        if code in _INLINE_SIZE:
            size = _INLINE_SIZE[code]
        else:
            (size,) = struct.unpack_from("<I", data, pos)
            pos += 4
        entries[key] = _decode(code, data[pos:pos + size])
        pos += size
    return entries, header
