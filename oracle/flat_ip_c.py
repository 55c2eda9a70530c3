"""ctypes wrapper of ``oracle/flat_ip.c`` (the C restatement of FAISS flat-IP search).
Test infrastructure and CPU baseline only — see the header of ``flat_ip.c``."""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_lib = None
_lib_kind = None


def build(native: bool = False) -> Path:
    target = "native" if native else "all"
    subprocess.run(["make", "-s", "-C", str(_DIR), target], check=True)
    return _DIR / "_build" / ("liboracle_flat_ip_native.so" if native else "liboracle_flat_ip.so")


def load(native: bool = False):
    """Load the portable build, or (``native=True``) a -march=native build made on this box."""
    global _lib, _lib_kind
    kind = "native" if native else "portable"
    if _lib is not None and _lib_kind == kind:
        return _lib
    try:
        path = build(native)
    except Exception:
        path = _DIR / "_build" / "liboracle_flat_ip.so"
        kind = "portable"
    L = C.CDLL(str(path))
    for name in ("oracle_flat_ip_search", "oracle_flat_ip_search_rowpar"):
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    L.oracle_set_blas.restype = C.c_int
    L.oracle_set_blas.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    L.oracle_has_blas.restype = C.c_int
    L.oracle_num_threads.restype = C.c_int
    L.oracle_set_num_threads.argtypes = [C.c_int]
    _lib, _lib_kind = L, kind
    return L


def try_attach_blas(L=None) -> str | None:
    """Point the blocked (nq >= 20) path at numpy's bundled OpenBLAS ``cblas_sgemm`` when it can
    be found; otherwise the portable register-blocked kernel in flat_ip.c is used."""
    L = L or load()
    libs_dir = Path(np.__file__).resolve().parent.parent / "numpy.libs"
    for path in sorted(glob.glob(str(libs_dir / "libscipy_openblas*.so*"))):
        for sym, ilp64 in (("scipy_cblas_sgemm64_", 1), ("scipy_cblas_sgemm", 0), ("cblas_sgemm64_", 1),
                           ("cblas_sgemm", 0)):
            if L.oracle_set_blas(path.encode(), sym.encode(), ilp64) == 0:
                return f"{os.path.basename(path)}:{sym}"
    return None


def flat_ip_search_c(xb, xq, k, *, rowpar: bool = False, native: bool = False):
    L = load(native)
    xb = np.ascontiguousarray(xb, dtype=np.float32)
    xq = np.ascontiguousarray(xq, dtype=np.float32)
    nq, d = xq.shape
    assert xb.shape[1] == d
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    fn = L.oracle_flat_ip_search_rowpar if rowpar else L.oracle_flat_ip_search
    rc = fn(xb.ctypes.data, xb.shape[0], d, xq.ctypes.data, nq, k, D.ctypes.data, I.ctypes.data)
    if rc != 0:
        raise RuntimeError("oracle_flat_ip_search: bad arguments (k must be > 0)")
    return D, I
