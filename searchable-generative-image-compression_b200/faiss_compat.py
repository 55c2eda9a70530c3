"""Drop-in for the seven ``faiss`` names the reference uses.

``IndexFlatIP(d)``, ``index.add``, ``index.search``, ``index.ntotal``, ``index.d``,
``read_index``, ``write_index`` — the call sites are src/search.py:69,76,115,
src/build.py:93-95,99,111,118,232-238 and src/compress.py:95,97,107,111 of the
reference.  Swap ``import faiss`` for ``from sgic_b200 import faiss_compat as faiss``
and those scripts run unchanged; the rows live in B200 HBM as fp16 (or bf16) and the
search runs in the hand-written sm_100a kernels behind ``libsgic.so``.

Argument checking follows the faiss Python wrapper: shape errors are
``AssertionError``, failures inside the library are ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, Sequence

import numpy as np

from . import _native
from ._native import SGIC_BF16, SGIC_F16, SGIC_RETAIN_F32, SGIC_RETAIN_U8, check

METRIC_INNER_PRODUCT = 0

_DTYPES = {"fp16": SGIC_F16, "float16": SGIC_F16, "f16": SGIC_F16, "half": SGIC_F16,
           "bf16": SGIC_BF16, "bfloat16": SGIC_BF16}

# rows*d*4 above which the host-side fp32 copy (kept so that write_index is bit-identical
# to faiss) is not created; the index file is then written from the 16-bit HBM rows.
RETAIN_FP32_MAX_BYTES = int(os.environ.get("SGIC_RETAIN_FP32_MAX_BYTES", str(8 << 30)))


def _default_device() -> int:
    env = os.environ.get("SGIC_DEVICE")
    if env is not None:
        return int(env)
    import sys
    torch = sys.modules.get("torch")
    if torch is not None and torch.cuda.is_available():
        return int(torch.cuda.current_device())
    lr = os.environ.get("LOCAL_RANK")
    return int(lr) if lr is not None else 0


def _dtype_code(dtype) -> int:
    if isinstance(dtype, int):
        return dtype
    key = str(dtype).replace("torch.", "").lower()
    if key not in _DTYPES:
        raise ValueError(f"dtype must be fp16 or bf16, got {dtype!r}")
    return _DTYPES[key]


def _torch_stream(device) -> C.c_void_p:
    """cudaStream_t of torch's current stream.  torch's default stream is the legacy stream
    (handle 0); the C ABI reads NULL as "the index's own stream", so pass cudaStreamLegacy."""
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream or 1)


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


class Index:
    """Base type (only used in annotations by the reference: src/search.py:65,113)."""

    d: int
    ntotal: int
    is_trained = True
    metric_type = METRIC_INNER_PRODUCT


class IndexFlatIP(Index):
    """Exact inner-product index resident in one B200's HBM.

    ``IndexFlatIP(d)`` is the faiss signature (src/build.py:93).  The keyword
    arguments are additive: ``dtype`` of the stored rows ("fp16" default, "bf16"),
    ``device`` ordinal, ``capacity`` rows to reserve, ``retain_fp32`` to keep the
    fp32 rows on the host for a bit-exact ``write_index``, ``retain_codes`` to keep the
    u8 ``clip_stream`` codes of rows added through ``add_u8`` / ``add_c2df`` (1 byte per
    element on the host) so that ``write_index`` emits exactly the fp32 rows the
    reference's ``dequantize_clip_u8`` would have given faiss (src/build.py:82-99).
    """

    def __init__(self, d: int, *, dtype="fp16", device: int | None = None, devices: Sequence[int] | None = None,
                 capacity: int = 0, retain_fp32: bool = True, retain_codes: bool = False, _handle=None,
                 _borrowed: bool = False):
        self._lib = _native.lib()
        self._h = C.c_void_p()
        self._borrowed = _borrowed
        if _handle is not None:
            self._h = _handle
            return
        flags = (SGIC_RETAIN_F32 if retain_fp32 else 0) | (SGIC_RETAIN_U8 if retain_codes else 0)
        if devices is not None:
            # ``devices=[0, 1, ...]``: the rows are sharded over these GPUs behind this one object, driven by this
            # one process (the reference's caller is one process: src/search.py:149-162); devices[0] is the home GPU
            devs = (C.c_int * len(devices))(*[int(x) for x in devices])
            check(self._lib.sgic_index_create_sharded(int(d), _dtype_code(dtype), len(devices), devs, int(capacity),
                                                      flags, C.byref(self._h)))
            return
        dev = _default_device() if device is None else int(device)
        check(self._lib.sgic_index_create(int(d), _dtype_code(dtype), dev, int(capacity), flags, C.byref(self._h)))

    # -- lifetime ---------------------------------------------------------------------
    def close(self) -> None:
        h, self._h = self._h, C.c_void_p()
        if h and not self._borrowed:
            self._lib.sgic_index_destroy(h)

    # -- shards (``devices=[...]``) --------------------------------------------------------
    @property
    def n_shards(self) -> int:
        return int(self._lib.sgic_index_n_shards(self._h))

    def shard(self, g: int) -> "IndexFlatIP":
        """The g-th GPU's rows as a borrowed single-GPU index (bulk loaders append to it directly and then call
        :meth:`adopt_shards`)."""
        h = C.c_void_p()
        check(self._lib.sgic_index_shard(self._h, int(g), C.byref(h)))
        return IndexFlatIP(0, _handle=h, _borrowed=True)

    def adopt_shards(self) -> None:
        """Make the rows the shards hold the index's rows: shard 0's first, then shard 1's, ..."""
        check(self._lib.sgic_index_adopt_shards(self._h))

    def save_shards(self, directory) -> None:
        """One ``shard-%05d-of-%05d.sgi2`` file per GPU (rows as stored in HBM)."""
        os.makedirs(str(directory), exist_ok=True)
        check(self._lib.sgic_index_save_shards(self._h, os.fsencode(str(directory))))

    def __del__(self):  # pragma: no cover - interpreter shutdown order
        try:
            self.close()
        except Exception:
            pass

    # -- faiss attributes -------------------------------------------------------------
    @property
    def d(self) -> int:
        return int(self._lib.sgic_index_d(self._h))

    @property
    def ntotal(self) -> int:
        return int(self._lib.sgic_index_ntotal(self._h))

    @property
    def dtype(self) -> str:
        return "bf16" if self._lib.sgic_index_dtype(self._h) == SGIC_BF16 else "fp16"

    @property
    def device(self) -> int:
        return int(self._lib.sgic_index_device(self._h))

    # -- faiss methods ----------------------------------------------------------------
    def add(self, x) -> None:
        """``index.add(X)`` — src/build.py:94, src/compress.py:107."""
        x = np.ascontiguousarray(x, dtype="float32")
        assert x.ndim == 2, "add expects a (n, d) array"
        n, d = x.shape
        assert d == self.d, f"add: got dimension {d}, index has {self.d}"
        if (self.ntotal + n) * d * 4 > RETAIN_FP32_MAX_BYTES:
            self._drop_retained()
        check(self._lib.sgic_index_add_f32(self._h, n, _ptr(x)))

    def search(self, x, k: int):
        """``D, I = index.search(q, k)`` — src/search.py:115."""
        x = np.ascontiguousarray(x, dtype="float32")
        assert x.ndim == 2, "search expects a (n, d) array"
        n, d = x.shape
        assert d == self.d, f"search: got dimension {d}, index has {self.d}"
        assert k > 0, "k must be positive"
        D = np.empty((n, k), dtype=np.float32)
        I = np.empty((n, k), dtype=np.int64)
        check(self._lib.sgic_index_search(self._h, n, _ptr(x), int(k), _ptr(D), _ptr(I)))
        return D, I

    def reset(self) -> None:
        check(self._lib.sgic_index_reset(self._h))

    def reconstruct_n(self, i0: int = 0, n: int | None = None) -> np.ndarray:
        n = self.ntotal - i0 if n is None else n
        out = np.empty((n, self.d), dtype=np.float32)
        check(self._lib.sgic_index_reconstruct(self._h, int(i0), int(n), _ptr(out)))
        return out

    def reconstruct(self, i: int) -> np.ndarray:
        return self.reconstruct_n(int(i), 1)[0]

    def codes(self, i0: int = 0, n: int | None = None) -> np.ndarray:
        """The retained u8 ``clip_stream`` codes of rows [i0, i0+n) (``retain_codes=True``)."""
        n = self.ntotal - i0 if n is None else n
        out = np.empty((n, self.d), dtype=np.uint8)
        check(self._lib.sgic_index_codes(self._h, int(i0), int(n), _ptr(out)))
        return out

    # -- additive surface (north-star items a / c) -----------------------------------------
    def reserve(self, rows: int) -> None:
        check(self._lib.sgic_index_reserve(self._h, int(rows)))

    def add_u8(self, q) -> None:
        """Append u8-quantised rows; dequantize_clip_u8 + l2n (src/search.py:16-22) run on device."""
        q = np.ascontiguousarray(q, dtype=np.uint8)
        assert q.ndim == 2 and q.shape[1] == self.d, "add_u8 expects (n, d) uint8"
        check(self._lib.sgic_index_add_u8(self._h, q.shape[0], _ptr(q)))

    def add_c2df(self, blob, offsets, n_threads: int = 0):
        """Batched ingest of ``.c2df`` files held back to back in ``blob``.

        Returns ``(n_added, status)`` with one status code per file (0 = added); failed
        files are skipped exactly as the ``[SKIP]`` branch of src/build.py:87-88 does.
        """
        blob = np.frombuffer(blob, dtype=np.uint8) if not isinstance(blob, np.ndarray) else blob
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = offsets.size - 1
        assert n >= 0 and (n == 0 or offsets[-1] <= blob.size), "offsets run past the blob"
        status = np.zeros(max(n, 0), dtype=np.int32)
        added = C.c_int64(0)
        check(self._lib.sgic_index_add_c2df(self._h, _ptr(blob), _ptr(offsets), n, _ptr(status),
                                            C.byref(added), int(n_threads)))
        return int(added.value), status

    def add_c2df_paths(self, paths: Sequence, n_threads: int = 0, batch_bytes: int = 256 << 20):
        """Read files from disk in batches and feed :meth:`add_c2df`. Returns per-path status."""
        statuses = []
        chunk, size = [], 0

        def flush():
            nonlocal chunk, size
            if not chunk:
                return
            offs = np.zeros(len(chunk) + 1, dtype=np.int64)
            np.cumsum([len(b) for b in chunk], out=offs[1:])
            _, st = self.add_c2df(np.frombuffer(b"".join(chunk), dtype=np.uint8), offs, n_threads)
            statuses.append(st)
            chunk, size = [], 0

        for p in paths:
            try:
                b = open(p, "rb").read()
            except OSError:
                b = b""  # unreadable -> bad magic -> skipped
            chunk.append(b)
            size += len(b)
            if size >= batch_bytes:
                flush()
        flush()
        return np.concatenate(statuses) if statuses else np.zeros(0, dtype=np.int32)

    def add_torch(self, x) -> None:
        """Append rows from a CUDA tensor on this index's device (fp32 → rounded; or already
        in the storage dtype → copied), on the current torch stream."""
        import torch
        assert x.is_cuda and x.dim() == 2 and x.shape[1] == self.d and x.is_contiguous()
        assert x.device.index == self.device
        st = _torch_stream(x.device)
        want = torch.bfloat16 if self.dtype == "bf16" else torch.float16
        if x.dtype == torch.float32:
            check(self._lib.sgic_index_add_f32_dev(self._h, x.shape[0], C.c_void_p(x.data_ptr()), st))
        elif x.dtype == want:
            check(self._lib.sgic_index_add_packed_dev(self._h, x.shape[0], C.c_void_p(x.data_ptr()), st))
        else:
            raise TypeError(f"add_torch: tensor dtype {x.dtype} does not match index dtype {self.dtype}")

    def search_torch(self, q, k: int, id_base: int = 0, out=None):
        """Device-resident search: ``q`` fp32 CUDA tensor (nq,d) → (D, I) CUDA tensors.
        Enqueued on the current torch stream; does not synchronise."""
        import torch
        assert q.is_cuda and q.dtype == torch.float32 and q.dim() == 2 and q.is_contiguous()
        assert q.shape[1] == self.d and q.device.index == self.device
        assert k > 0
        nq = q.shape[0]
        if out is None:
            D = torch.empty((nq, k), dtype=torch.float32, device=q.device)
            I = torch.empty((nq, k), dtype=torch.int64, device=q.device)
        else:
            D, I = out
        st = _torch_stream(q.device)
        check(self._lib.sgic_index_search_dev(self._h, nq, C.c_void_p(q.data_ptr()), int(k),
                                              C.c_void_p(D.data_ptr()), C.c_void_p(I.data_ptr()),
                                              int(id_base), st))
        return D, I

    def set_option(self, name: str, value: int) -> None:
        check(self._lib.sgic_index_set_option(self._h, name.encode(), int(value)))

    def scan_times_ms(self, max_n: int = 256) -> np.ndarray:
        """Durations of the scan kernels of the last searches made under ``set_option("timing", 2)`` (CUDA events
        on the search's stream, recorded without a synchronise)."""
        out = np.zeros(max_n, dtype=np.float32)
        n = C.c_int(0)
        check(self._lib.sgic_index_scan_times(self._h, int(max_n), _ptr(out), C.byref(n)))
        return out[:n.value].copy()

    def stat(self, name: str) -> int:
        return int(self._lib.sgic_index_get_stat(self._h, name.encode()))

    def _drop_retained(self) -> None:
        # an add through any non-fp32-host path clears the retained copy inside the library;
        # here we only need to stop retaining: reset flag by a zero-row device add
        if self.stat("retained_rows") >= 0:
            check(self._lib.sgic_index_set_option(self._h, b"drop_retained", 1))


def write_shard(index: IndexFlatIP, path: str, *, row_start: int = 0, total_rows: int = -1, shard: int = 0,
                n_shards: int = 1) -> None:
    """Additive: write the rows as they sit in HBM (fp16 / bf16) in the "SGI2" layout —
    half the bytes of ``write_index`` and no conversion on load.  ``read_index`` loads either format."""
    check(index._lib.sgic_index_write_v2(index._h, os.fsencode(str(path)), int(row_start), int(total_rows),
                                         int(shard), int(n_shards)))


def shard_info(index: IndexFlatIP) -> dict:
    """Placement recorded in the SGI2 file an index was loaded from."""
    out = (C.c_int64 * 4)()
    check(index._lib.sgic_index_shard_info(index._h, out))
    return {"row_start": int(out[0]), "total_rows": int(out[1]), "shard": int(out[2]), "n_shards": int(out[3])}


def codes_to_f32(q) -> np.ndarray:
    """``dequantize_clip_u8`` + ``l2n`` (src/search.py:16-22) for a block of u8 rows on the host, with numpy's
    operation order — bit-identical to what the reference computes row by row."""
    q = np.ascontiguousarray(q, dtype=np.uint8)
    assert q.ndim == 2
    out = np.empty(q.shape, dtype=np.float32)
    check(_native.lib().sgic_codes_to_f32(_ptr(q), q.shape[0], q.shape[1], _ptr(out)))
    return out


def read_index_shards(directory, devices: Sequence[int], *, retain_fp32: bool = False) -> IndexFlatIP:
    """Load a directory of ``shard-*-of-*.sgi2`` files (``IndexFlatIP.save_shards`` / ``ShardedIndexFlatIP.save``)
    onto ``devices``, one file per GPU, behind one index object."""
    lib = _native.lib()
    h = C.c_void_p()
    devs = (C.c_int * len(devices))(*[int(x) for x in devices])
    check(lib.sgic_index_load_shards(os.fsencode(str(directory)), len(devices), devs,
                                     SGIC_RETAIN_F32 if retain_fp32 else 0, C.byref(h)))
    return IndexFlatIP(0, _handle=h)


def write_index(index: IndexFlatIP, path: str) -> None:
    """``faiss.write_index(index, path)`` — src/build.py:95,99; src/compress.py:111."""
    check(index._lib.sgic_index_write(index._h, os.fsencode(str(path))))


def read_index(path: str, *, dtype="fp16", device: int | None = None, retain_fp32: bool | None = None) -> IndexFlatIP:
    """``faiss.read_index(path)`` — src/search.py:69,76; src/compress.py:95."""
    lib = _native.lib()
    h = C.c_void_p()
    if retain_fp32 is None:
        try:
            retain_fp32 = os.path.getsize(path) <= RETAIN_FP32_MAX_BYTES
        except OSError:
            retain_fp32 = True
    flags = SGIC_RETAIN_F32 if retain_fp32 else 0
    dev = _default_device() if device is None else int(device)
    check(lib.sgic_index_read(os.fsencode(str(path)), _dtype_code(dtype), dev, flags, C.byref(h)))
    return IndexFlatIP(0, _handle=h)
