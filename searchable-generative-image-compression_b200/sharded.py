"""Row-sharded index over the GPUs of one box (north-star item (c)).

One process per GPU (``torchrun``); every rank holds a slice of the rows in its own HBM as a
:class:`faiss_compat.IndexFlatIP`.  A search is: queries replicated on every rank → local
top-k on each GPU (K3/K4) with global row numbers → ONE ``all_gather`` of the packed
``(nq, k)`` candidates (NCCL over NVLink; 12·nq·k bytes per rank) → on-device merge
``G·k → k`` (K5, ``sgic_merge_topk_dev``) ordered by (score desc, global id asc), so the
G-GPU answer is identical to the 1-GPU answer.  There is no other collective on the path.

When the ranks can map each other's memory (CUDA IPC, one box) the exchange runs without a
collective call at all (K5x, ``sgic_xchg_merge_dev``): every rank stores its candidates straight
into the peers' HBM over NVLink, raises a flag, and merges its own buffer as soon as all flags
are up — two small launches instead of pack + all_gather + unpack + merge.  NCCL stays the
transport for anything the buffer does not hold (``nq·k`` above ``PEER_MAX_CANDS``) and when
``SGIC_EXCHANGE=nccl`` asks for it.

The reference has no multi-GPU retrieval (its only ``torch.distributed`` use shards the image
encoder over files, src/compress.py:34-55,293-306, with rank 0 building the index serially);
this class is the additive surface the north star asks for and keeps the faiss names.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Tuple

import numpy as np

from . import _native
from .faiss_compat import Index, IndexFlatIP, _torch_stream


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range of ``rank`` when ``n`` rows are split over ``world`` ranks:
    ``[rank*ceil(n/world), min(n, (rank+1)*ceil(n/world)))``."""
    per = -(-n // world) if n > 0 else 0
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


PEER_MAX_CANDS = 1 << 19   # nq*k the peer-exchange buffer holds per rank (12 MB per rank and parity at 8 GPUs)


class PeerExchange:
    """K5x: the per-rank exchange buffer, its IPC handle swap, and the push + wait + merge call."""

    def __init__(self, dist, group, device_index: int, world: int, rank: int, max_cands: int = PEER_MAX_CANDS):
        import torch
        lib = _native.lib()
        self._lib, self._h, self.max_cands = lib, C.c_void_p(), int(max_cands)
        self._dist, self._group = dist, group
        ok, handle = True, bytes(64)
        try:
            _native.check(lib.sgic_xchg_create(device_index, world, rank, self.max_cands, C.byref(self._h)))
            buf = (C.c_uint8 * 64)()
            _native.check(lib.sgic_xchg_export(self._h, buf))
            handle = bytes(buf)
        except RuntimeError:
            ok = False
        gathered = [None] * world
        dist.all_gather_object(gathered, (ok, handle), group=group)
        if ok and all(g[0] for g in gathered):
            try:
                blob = b"".join(g[1] for g in gathered)
                _native.check(lib.sgic_xchg_open(self._h, C.c_char_p(blob)))
            except RuntimeError:
                ok = False
        else:
            ok = False
        verdict = [None] * world
        dist.all_gather_object(verdict, ok, group=group)   # every rank mapped every peer, or nobody uses it
        self.ok = all(verdict)
        if not self.ok:
            self.close(collective=False)
        torch.cuda.synchronize()
        dist.barrier(group=group)

    def merge(self, D_local, I_local, k: int, by_position: bool):
        import torch
        nq = D_local.shape[0]
        D = torch.empty((nq, k), dtype=torch.float32, device=D_local.device)
        I = torch.empty((nq, k), dtype=torch.int64, device=D_local.device)
        _native.check(self._lib.sgic_xchg_merge_dev(self._h, nq, k, C.c_void_p(D_local.data_ptr()),
                                                    C.c_void_p(I_local.data_ptr()), C.c_void_p(D.data_ptr()),
                                                    C.c_void_p(I.data_ptr()), 1 if by_position else 0,
                                                    _torch_stream(D_local.device)))
        return D, I

    def search_host(self, local: IndexFlatIP, x: np.ndarray, k: int, id_base: int, by_position: bool):
        """The whole step in one C call: staging, local scan, exchange + merge, answer back on the host."""
        nq = x.shape[0]
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
        _native.check(self._lib.sgic_xchg_search(self._h, local._h, nq, C.c_void_p(x.ctypes.data), k,
                                                 C.c_void_p(D.ctypes.data), C.c_void_p(I.ctypes.data), int(id_base),
                                                 1 if by_position else 0))
        return D, I

    def check(self) -> None:
        _native.check(self._lib.sgic_xchg_error(self._h))

    def close(self, collective: bool = True) -> None:
        if self._h:
            if collective:
                import torch
                torch.cuda.synchronize()
                self._dist.barrier(group=self._group)   # nobody still writes into a buffer that is about to go
            self._lib.sgic_xchg_destroy(self._h)
            self._h = C.c_void_p()


def _merge_cuda(D_lists, I_lists, k: int, by_position: bool):
    """K5 through the C ABI.  ``D_lists``/``I_lists``: (G, nq, k) CUDA tensors."""
    import torch
    G, nq, _ = D_lists.shape
    D = torch.empty((nq, k), dtype=torch.float32, device=D_lists.device)
    I = torch.empty((nq, k), dtype=torch.int64, device=D_lists.device)
    lib = _native.lib()
    _native.check(lib.sgic_merge_topk_dev(D_lists.device.index, nq, G, k,
                                          C.c_void_p(D_lists.data_ptr()), C.c_void_p(I_lists.data_ptr()),
                                          C.c_void_p(D.data_ptr()), C.c_void_p(I.data_ptr()),
                                          1 if by_position else 0, _torch_stream(D_lists.device)))
    return D, I


class ShardedIndexFlatIP(Index):
    """``IndexFlatIP`` whose rows are spread over the ranks of a ``torch.distributed`` group.

    Every method is collective: all ranks call it with the same arguments (``add`` /
    ``search`` take the same arrays on every rank, as replicated queries do).
    ``local_factory`` / ``merge_fn`` exist so that the host logic can be exercised on CPU
    (gloo) with stand-ins; the product path uses the CUDA defaults.
    """

    def __init__(self, d: int, *, dtype="fp16", group=None, device: Optional[int] = None,
                 local_factory: Optional[Callable] = None, merge_fn: Optional[Callable] = None):
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("ShardedIndexFlatIP needs torch.distributed to be initialised (torchrun)")
        self._dist = dist
        self._group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._d = int(d)
        self._cuda = dist.get_backend(group) == "nccl"
        if self._cuda:
            dev = torch.cuda.current_device() if device is None else int(device)
            self._dev = torch.device("cuda", dev)
        else:
            dev = None
            self._dev = torch.device("cpu")
        self.local = local_factory(d) if local_factory else IndexFlatIP(d, dtype=dtype, device=dev, retain_fp32=False)
        self._merge = merge_fn or _merge_cuda
        self._pin_q = self._pin_D = self._pin_I = None
        self._peer = None
        import os
        if (self._cuda and self.world > 1 and merge_fn is None
                and os.environ.get("SGIC_EXCHANGE", "peer") != "nccl" and self.world <= 16):
            px = PeerExchange(dist, group, dev, self.world, self.rank)
            self._peer = px if px.ok else None
        # segments of this rank: (global_start, local_start, count), ascending in both
        self._segments: List[Tuple[int, int, int]] = []
        self._ntotal = 0

    # ---------------------------------------------------------------- faiss attributes
    @property
    def d(self) -> int:
        return self._d

    @property
    def ntotal(self) -> int:
        return self._ntotal

    @property
    def local_ntotal(self) -> int:
        return self.local.ntotal

    # ---------------------------------------------------------------- add
    def _note_segment(self, global_start: int, count: int) -> None:
        if count <= 0:
            return
        local_start = self.local.ntotal - count
        if self._segments and self._segments[-1][0] + self._segments[-1][2] == global_start:
            g0, l0, c0 = self._segments[-1]
            self._segments[-1] = (g0, l0, c0 + count)
        else:
            self._segments.append((global_start, local_start, count))

    def add(self, x) -> None:
        """Same ``x`` (n, d) on every rank; rank g keeps the g-th contiguous slice."""
        x = np.ascontiguousarray(x, dtype="float32")
        assert x.ndim == 2 and x.shape[1] == self._d
        lo, hi = shard_range(x.shape[0], self.world, self.rank)
        if hi > lo:
            self.local.add(x[lo:hi])
        self._note_segment(self._ntotal + lo, hi - lo)
        self._ntotal += x.shape[0]

    def add_local(self, adder: Callable[[IndexFlatIP], None], global_start: int, count: int, total: int) -> None:
        """Loader path: ``adder(local_index)`` appends this rank's ``count`` rows (by any of the
        local index's add_* methods); they are global rows ``[global_start, global_start+count)``
        of a block of ``total`` rows appended collectively."""
        before = self.local.ntotal
        adder(self.local)
        assert self.local.ntotal - before == count, "adder appended a different number of rows"
        self._note_segment(self._ntotal + global_start, count)
        self._ntotal += total

    def add_c2df_paths(self, paths, n_threads: int = 0):
        """Collective ingest of ``.c2df`` files (config C5: header ingestion straight to the owning GPU).

        Every rank passes the SAME path list (``sorted(glob)`` order, as build.py:74 produces); rank g reads,
        walks and decodes the g-th contiguous slice of it on its own host threads and GPU
        (``IndexFlatIP.add_c2df_paths``).  Files that fail are skipped as ``[SKIP]`` does (build.py:87-88), so the
        global row number of a file is the number of good files before it: one ``all_gather`` of the per-rank
        counts (and statuses) places every shard.  Returns the status of every file, identical on all ranks."""
        paths = list(paths)
        lo, hi = shard_range(len(paths), self.world, self.rank)
        before = self.local.ntotal
        st_local = np.asarray(self.local.add_c2df_paths(paths[lo:hi], n_threads), dtype=np.int32)
        count = self.local.ntotal - before
        gathered = [None] * self.world
        self._dist.all_gather_object(gathered, (int(count), st_local.tolist()), group=self._group)
        start = sum(c for c, _ in gathered[:self.rank])
        total = sum(c for c, _ in gathered)
        self._note_segment(self._ntotal + start, count)
        self._ntotal += total
        return np.concatenate([np.asarray(st, dtype=np.int32) for _, st in gathered]) if paths else np.zeros(0, np.int32)

    # ---------------------------------------------------------------- search
    def _to_global(self, I_local):
        """local row numbers → global row numbers (identity + base for a single segment)."""
        import torch
        if len(self._segments) <= 1:
            base = self._segments[0][0] - self._segments[0][1] if self._segments else 0
            return torch.where(I_local >= 0, I_local + base, I_local) if base else I_local
        l_starts = torch.tensor([s[1] for s in self._segments], dtype=torch.int64, device=I_local.device)
        offs = torch.tensor([s[0] - s[1] for s in self._segments], dtype=torch.int64, device=I_local.device)
        seg = torch.bucketize(I_local.clamp(min=0), l_starts, right=True) - 1
        return torch.where(I_local >= 0, I_local + offs[seg], I_local)

    def search_torch(self, q, k: int):
        """``q``: fp32 (nq, d) tensor on this rank's device, identical on every rank.
        Returns device tensors (D, I), identical on every rank."""
        import torch
        assert k > 0
        nq = q.shape[0]
        if len(self._segments) == 1 and self._cuda:
            base = self._segments[0][0] - self._segments[0][1]
            D, I = self.local.search_torch(q, k, id_base=base)        # id_base folded into the merge kernel
        else:
            D, I = self.local.search_torch(q, k)
            I = self._to_global(I)
        if self.world == 1:
            return D, I
        contiguous_shards = self._ntotal >= (1 << 32)   # ids beyond 32 bits: tie-break by shard position
        if self._peer is not None and nq * k <= self._peer.max_cands:
            return self._peer.merge(D.contiguous(), I.contiguous(), k, contiguous_shards)
        # pack (ids, scores) into one buffer so that the exchange is a single all_gather
        packed = torch.empty((nq, k, 3), dtype=torch.int32, device=q.device)
        packed[..., :2] = I.view(torch.int32).view(nq, k, 2)
        packed[..., 2] = D.view(torch.int32)
        gathered = torch.empty((self.world * nq, k, 3), dtype=torch.int32, device=q.device)
        self._dist.all_gather_into_tensor(gathered, packed, group=self._group)
        gathered = gathered.view(self.world, nq, k, 3)
        I_lists = gathered[..., :2].contiguous().view(torch.int64).view(self.world, nq, k)
        D_lists = gathered[..., 2].contiguous().view(torch.float32)
        return self._merge(D_lists, I_lists, k, contiguous_shards)

    @property
    def exchange(self) -> str:
        """"peer" (K5x: stores into the peers' HBM + flags) or "nccl" (all_gather + K5)."""
        return "peer" if self._peer is not None else "nccl"

    def close(self) -> None:
        """Collective.  Releases the peer-exchange buffer (after a barrier) and the local shard."""
        if self._peer is not None:
            self._peer.check()
            self._peer.close()
            self._peer = None
        if hasattr(self.local, "close"):
            self.local.close()

    def search(self, x, k: int):
        """faiss signature: host fp32 (nq, d) in, host (D, I) out — on every rank."""
        import torch
        x = np.ascontiguousarray(x, dtype="float32")
        assert x.ndim == 2 and x.shape[1] == self._d, "search expects (n, d)"
        assert k > 0, "k must be positive"
        if not self._cuda:
            D, I = self.search_torch(torch.from_numpy(x).to(self._dev), k)
            return D.cpu().numpy(), I.cpu().numpy()
        nq = x.shape[0]
        if (self._peer is not None and nq * k <= self._peer.max_cands and len(self._segments) == 1
                and isinstance(self.local, IndexFlatIP)):
            base = self._segments[0][0] - self._segments[0][1]
            return self._peer.search_host(self.local, x, k, base, self._ntotal >= (1 << 32))
        # pinned staging both ways, one synchronisation per call (what sgic_index_search does on one GPU)
        need_in, need_out = nq * self._d, nq * k
        if self._pin_q is None or self._pin_q.numel() < need_in:
            self._pin_q = torch.empty(max(need_in, 1 << 16), dtype=torch.float32).pin_memory()
        if self._pin_D is None or self._pin_D.numel() < need_out:
            self._pin_D = torch.empty(max(need_out, 1 << 14), dtype=torch.float32).pin_memory()
            self._pin_I = torch.empty(max(need_out, 1 << 14), dtype=torch.int64).pin_memory()
        pq = self._pin_q[:need_in].view(nq, self._d)
        pq.copy_(torch.from_numpy(x))
        q = pq.to(self._dev, non_blocking=True)
        D, I = self.search_torch(q, k)
        pD, pI = self._pin_D[:need_out].view(nq, k), self._pin_I[:need_out].view(nq, k)
        pD.copy_(D, non_blocking=True)
        pI.copy_(I, non_blocking=True)
        torch.cuda.current_stream(self._dev).synchronize()
        return pD.numpy().copy(), pI.numpy().copy()

    # ---------------------------------------------------------------- persistence (SGI2 shard files)
    @staticmethod
    def shard_path(directory, rank: int, world: int) -> str:
        import os
        return os.path.join(str(directory), f"shard-{rank:05d}-of-{world:05d}.sgi2")

    def save(self, directory) -> None:
        """Collective.  Every rank writes the rows it holds, as they sit in HBM, to its own SGI2 file
        (``faiss_compat.write_shard``); the header records where the rows sit in the logical index."""
        import os
        from .faiss_compat import write_shard
        if len(self._segments) > 1:
            raise RuntimeError("save() needs one contiguous row range per rank (build with add_local / one add)")
        os.makedirs(str(directory), exist_ok=True)
        start = self._segments[0][0] if self._segments else 0
        write_shard(self.local, self.shard_path(directory, self.rank, self.world), row_start=start,
                    total_rows=self._ntotal, shard=self.rank, n_shards=self.world)
        if self.world > 1:
            self._dist.barrier(group=self._group)

    @classmethod
    def load(cls, directory, *, group=None, device: Optional[int] = None) -> "ShardedIndexFlatIP":
        """Collective.  Rank g loads ``shard-g-of-G.sgi2`` straight into its HBM (no conversion)."""
        import torch.distributed as dist
        from .faiss_compat import read_index, shard_info
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        local = read_index(cls.shard_path(directory, rank, world), device=device, retain_fp32=False)
        info = shard_info(local)
        if info["n_shards"] != world or info["shard"] != rank:
            raise RuntimeError(f"{directory}: shard file written for {info['n_shards']} ranks, loaded with {world}")
        self = cls(local.d, dtype=local.dtype, group=group, device=device, local_factory=lambda d: local)
        self._segments = [(info["row_start"], 0, local.ntotal)] if local.ntotal else []
        self._ntotal = info["total_rows"]
        return self
