"""ctypes binding of the C-ABI library ``libsgic.so`` (include/sgic.h).

There is no CPU fallback: if the library is missing, cannot be loaded, or no B200 is
visible when an index is created, the caller gets an exception that says so.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = _PKG_DIR / "libsgic.so"

_lib = None

c_i64_p = C.POINTER(C.c_int64)
c_i32_p = C.POINTER(C.c_int32)
c_f32_p = C.POINTER(C.c_float)
c_u8_p = C.POINTER(C.c_uint8)

# name -> (restype, argtypes); mirrors include/sgic.h one to one
_SIGNATURES = {
    "sgic_last_error": (C.c_char_p, []),
    "sgic_version": (C.c_int, []),
    "sgic_index_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_void_p)]),
    "sgic_index_destroy": (C.c_int, [C.c_void_p]),
    "sgic_index_create_sharded": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int64, C.c_int,
                                            C.POINTER(C.c_void_p)]),
    "sgic_index_n_shards": (C.c_int, [C.c_void_p]),
    "sgic_index_shard": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "sgic_index_adopt_shards": (C.c_int, [C.c_void_p]),
    "sgic_index_save_shards": (C.c_int, [C.c_void_p, C.c_char_p]),
    "sgic_index_load_shards": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    "sgic_index_ntotal": (C.c_int64, [C.c_void_p]),
    "sgic_index_d": (C.c_int, [C.c_void_p]),
    "sgic_index_dtype": (C.c_int, [C.c_void_p]),
    "sgic_index_device": (C.c_int, [C.c_void_p]),
    "sgic_index_reserve": (C.c_int, [C.c_void_p, C.c_int64]),
    "sgic_index_reset": (C.c_int, [C.c_void_p]),
    "sgic_index_add_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    "sgic_index_add_f32_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "sgic_index_add_packed_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "sgic_index_add_u8": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    "sgic_index_add_u8_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "sgic_index_add_c2df": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                      C.POINTER(C.c_int64), C.c_int]),
    "sgic_c2df_parse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_int]),
    "sgic_index_search": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "sgic_index_search_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_int64, C.c_void_p]),
    "sgic_merge_topk_dev": (C.c_int, [C.c_int, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "sgic_xchg_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_void_p)]),
    "sgic_xchg_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sgic_xchg_open": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sgic_xchg_merge_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int, C.c_void_p]),
    "sgic_xchg_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_int64, C.c_int]),
    "sgic_xchg_error": (C.c_int, [C.c_void_p]),
    "sgic_xchg_destroy": (C.c_int, [C.c_void_p]),
    "sgic_index_write": (C.c_int, [C.c_void_p, C.c_char_p]),
    "sgic_index_read": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "sgic_index_write_v2": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64, C.c_int64, C.c_int, C.c_int]),
    "sgic_index_shard_info": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sgic_index_reconstruct": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "sgic_index_codes": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "sgic_codes_to_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "sgic_index_data_dev": (C.c_void_p, [C.c_void_p]),
    "sgic_index_scan_times": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]),
    "sgic_index_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "sgic_index_get_stat": (C.c_int64, [C.c_void_p, C.c_char_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

SGIC_F16, SGIC_BF16 = 0, 1
SGIC_RETAIN_F32 = 1
SGIC_RETAIN_U8 = 2

C2DF_STATUS = {
    0: "ok",
    1: "bad magic",
    2: "truncated container",
    3: "No 'clip_stream' or 'clip_meta' was found, this file can't be used to search!",
    4: "Invalid clip_meta.dim",
    5: "zstd decode error",
    6: "Dimension didn't match",
    7: "vector dimension differs from the index dimension",
    8: "unknown type code",
    9: "the header or an entry does not load (invalid JSON / UTF-8 / array entry)",
}


class NativeLibraryError(ImportError):
    pass


def lib():
    """Load ``libsgic.so`` once; raise loudly when it is not there (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("SGIC_LIB", str(LIB_PATH)))
    if not path.exists():
        raise NativeLibraryError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). This package has no CPU / PyTorch fallback.")
    try:
        L = C.CDLL(str(path))
    except OSError as e:  # pragma: no cover - depends on the box
        raise NativeLibraryError(f"could not load {path}: {e}") from e
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def last_error() -> str:
    msg = lib().sgic_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int) -> None:
    """Map a non-zero status to RuntimeError (how faiss surfaces its C++ exceptions)."""
    if rc != 0:
        raise RuntimeError(last_error() or f"libsgic call failed with status {rc}")
