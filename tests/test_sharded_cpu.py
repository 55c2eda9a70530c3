"""Host logic of the sharded index on CPU: world_size-2 gloo processes, with the local GPU
index and the K5 merge replaced by oracle-backed stand-ins (tests only — the product classes
take them as injectable callables and default to the CUDA implementations)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def test_shard_range_partition():
    from sgic_b200.sharded import shard_range
    for n in (0, 1, 7, 8, 100, 100_000_000, 1_000_000_007):
        for w in (1, 2, 4, 8):
            parts = [shard_range(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            for (a0, a1), (b0, b1) in zip(parts, parts[1:]):
                assert a1 == b0 and a0 <= a1
            per = -(-n // w) if n else 0
            assert all(hi - lo <= per for lo, hi in parts)
    assert shard_range(100_000_000, 8, 3) == (37_500_000, 50_000_000)


class _FakeLocal:
    """Oracle-backed stand-in for faiss_compat.IndexFlatIP (CPU, torch tensors in/out)."""

    def __init__(self, d):
        self.d = d
        self.rows = np.zeros((0, d), dtype=np.float32)

    @property
    def ntotal(self):
        return self.rows.shape[0]

    def add(self, x):
        self.rows = np.concatenate([self.rows, np.asarray(x, dtype=np.float32)])

    def add_c2df_paths(self, paths, n_threads=0):
        """oracle decode of each file; status 0 = added, 1 = skipped (any reference exception)"""
        from oracle import c2df_ref
        st = []
        for p in paths:
            try:
                _, z = c2df_ref.decode_clip(Path(p).read_bytes())
                if z.shape[0] != self.d:
                    raise ValueError("dim")
                self.add(z[None, :])
                st.append(0)
            except Exception:
                st.append(1)
        return np.asarray(st, dtype=np.int32)

    def search_torch(self, q, k, id_base=0):
        from oracle.flat_ip import flat_ip_search
        D, I = flat_ip_search(self.rows, q.numpy(), k)
        I = np.where(I >= 0, I + id_base, I)
        return torch.from_numpy(D), torch.from_numpy(I)


def _fake_merge(D_lists, I_lists, k, by_position):
    """(score desc, global id asc) merge of the gathered lists — what K5 computes."""
    G, nq, _ = D_lists.shape
    D = np.full((nq, k), -3.4028234663852886e38, dtype=np.float32)
    I = np.full((nq, k), -1, dtype=np.int64)
    Dl, Il = D_lists.numpy(), I_lists.numpy()
    for qi in range(nq):
        d = Dl[:, qi, :].reshape(-1)
        i = Il[:, qi, :].reshape(-1)
        valid = np.nonzero(i >= 0)[0]
        order = valid[np.lexsort((i[valid], -d[valid].astype(np.float64)))][:k]
        D[qi, :order.size] = d[order]
        I[qi, :order.size] = i[order]
    return torch.from_numpy(D), torch.from_numpy(I)


def _worker(rank, world, port, seed, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sgic_b200.sharded import ShardedIndexFlatIP
        from oracle.flat_ip import flat_ip_search
        rng = np.random.default_rng(seed)                     # same data on every rank
        d, k = 64, 10
        base = rng.standard_normal((301, d)).astype(np.float32)
        base /= np.linalg.norm(base, axis=1, keepdims=True)
        blocks = [base[:120], base[120:121], base[121:], base[:50]]   # ragged adds; the last one duplicates rows
        full = np.concatenate(blocks)
        idx = ShardedIndexFlatIP(d, local_factory=_FakeLocal, merge_fn=_fake_merge)
        for b in blocks:
            idx.add(b)
        assert idx.ntotal == full.shape[0] and idx.d == d
        assert idx.exchange == "nccl"      # the peer exchange (K5x) needs CUDA IPC: never on a CPU / gloo group
        totals = [None] * world
        dist.all_gather_object(totals, idx.local_ntotal)
        assert sum(totals) == full.shape[0]
        q = full[[3, 200, 45]] + 0.01 * rng.standard_normal((3, d)).astype(np.float32)
        D, I = idx.search(q, k)
        Dref, Iref = flat_ip_search(full, q, k)
        assert np.array_equal(I, Iref), (rank, I, Iref)         # ties (duplicates) resolve to the lowest global id
        assert np.allclose(D, Dref, atol=1e-6)
        D2, I2 = idx.search(q[:1], 400)                          # k > ntotal: -1 padding survives the merge
        assert (I2[0] >= 0).sum() == full.shape[0] and np.all(I2[0, full.shape[0]:] == -1)
        # collective .c2df ingest: every rank takes its slice of the sorted path list; skipped files shift the
        # global row numbers of everything behind them, on every rank alike
        paths = sorted(str(p) for p in Path(out_dir).glob("corpus/*.c2df"))
        idx2 = ShardedIndexFlatIP(d, local_factory=_FakeLocal, merge_fn=_fake_merge)
        status = idx2.add_c2df_paths(paths)
        good = [i for i, p in enumerate(paths) if "bad" not in Path(p).name]
        assert list(np.nonzero(status == 0)[0]) == good and idx2.ntotal == len(good)
        from oracle import c2df_ref
        rows = np.stack([c2df_ref.decode_clip(Path(paths[i]).read_bytes())[1] for i in good])
        D3, I3 = idx2.search(rows[[0, len(good) - 1]], 5)
        Dr, Ir = flat_ip_search(rows, rows[[0, len(good) - 1]], 5)
        assert np.array_equal(I3, Ir) and I3[0, 0] == 0 and I3[1, 0] == len(good) - 1
        idx.close()                        # collective; with stand-ins there is nothing to release
        idx2.close()
        (Path(out_dir) / f"ok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


def _write_c2df_corpus(root, d=64, n=23):
    from sgic_b200 import c2df
    from sgic_b200.index_build import quantize_u8_and_compress
    rng = np.random.default_rng(5)
    (root / "corpus").mkdir()
    for i in range(n):
        v = rng.standard_normal(d).astype(np.float32)
        v /= np.linalg.norm(v)
        payload, meta = quantize_u8_and_compress(v)
        name = f"img_{i:03d}.c2df"
        blob = c2df.pack_c2df({"clip_stream": payload, "clip_meta": meta}, {"version": 2})
        if i in (2, 11, 12):                                   # broken files land in both ranks' slices
            name, blob = f"img_{i:03d}_bad.c2df", b"XXXX" + blob[4:]
        (root / "corpus" / name).write_bytes(blob)


def test_two_rank_gloo_matches_single_index(tmp_path):
    _write_c2df_corpus(tmp_path)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, 123, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
