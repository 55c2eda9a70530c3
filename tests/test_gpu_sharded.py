"""K5 (merge of per-shard answers) and the sharded index on real GPUs."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def unit(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


@pytest.mark.parametrize("G,nq,k,by_pos", [(2, 1, 10, False), (8, 5, 10, False), (8, 3, 100, True), (3, 2, 1024, False),
                                           (16, 1, 1024, True)])
def test_k5_merge_equals_single_index(G, nq, k, by_pos):
    """G shards emulated on one GPU (one kernel over all ranks' lists — no cross-launch waiting):
    merged answer must be IDENTICAL (ids and scores) to one index over all rows."""
    import torch
    from sgic_b200 import faiss_compat as faiss
    from sgic_b200.sharded import _merge_cuda, shard_range
    rng = np.random.default_rng(G * 100 + nq + k)
    n, d = 6000, 512
    base = unit(rng, n // 2, d)
    xb = np.concatenate([base, base])[rng.permutation(n)] if not by_pos else np.concatenate([base, base])
    xq = unit(rng, nq, d)
    whole = faiss.IndexFlatIP(d, device=0)
    whole.add(xb)
    Dw, Iw = whole.search(xq, k)
    q = torch.from_numpy(xq).cuda()
    Dl, Il = [], []
    for g in range(G):
        lo, hi = shard_range(n, G, g)
        sh = faiss.IndexFlatIP(d, device=0)
        sh.add(xb[lo:hi])
        D, I = sh.search_torch(q, k, id_base=lo)
        Dl.append(D)
        Il.append(I)
    D, I = _merge_cuda(torch.stack(Dl), torch.stack(Il), k, by_pos)
    torch.cuda.synchronize()
    assert np.array_equal(I.cpu().numpy(), Iw)
    assert np.array_equal(D.cpu().numpy(), Dw)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from sgic_b200.sharded import ShardedIndexFlatIP
        from oracle.flat_ip import check_topk
        rng = np.random.default_rng(77)
        d = 512
        xb = unit(rng, 50_001, d)
        xb[40_000:40_100] = xb[:100]                       # duplicates across shards
        idx = ShardedIndexFlatIP(d)
        idx.add(xb[:30_000])
        idx.add(xb[30_000:])                               # two segments per rank
        assert idx.ntotal == xb.shape[0]
        # the exchange is K5x (stores into the peer's HBM over NVLink + flags) wherever the ranks can map each
        # other's memory; the NCCL all_gather + K5 route must give bit-identical answers
        os.environ["SGIC_EXCHANGE"] = "nccl"
        idx_nccl = ShardedIndexFlatIP(d)
        del os.environ["SGIC_EXCHANGE"]
        idx_nccl.add(xb[:30_000])
        idx_nccl.add(xb[30_000:])
        assert idx_nccl.exchange == "nccl"
        (Path(out_dir) / f"exchange{rank}").write_text(idx.exchange)
        for nq, k in ((1, 10), (6, 10), (2, 100), (300, 10), (1, 1), (3, 1024)):
            xq = np.concatenate([xb[:1], unit(rng, nq - 1, d)]) if nq > 1 else xb[:1].copy()
            for rep in range(3):                               # both parities of the exchange buffer, and a reuse
                D, I = idx.search(xq, k)
            Dn, In = idx_nccl.search(xq, k)
            assert np.array_equal(I, In) and np.array_equal(D, Dn)
            xb16 = xb.astype(np.float16).astype(np.float64)
            sel = slice(0, min(nq, 8))
            check_topk(D[sel], I[sel], xb16, xq[sel].astype(np.float16).astype(np.float64), k, score_tol=2e-5, tie_tol=1e-6)
            if k >= 2:
                assert I[0, 0] == 0 and I[0, 1] == 40_000      # exact tie -> lower global id first
        # candidates beyond the exchange buffer (nq * k > PEER_MAX_CANDS) travel by all_gather
        xq = unit(rng, 600, d)
        D, I = idx.search(xq, 1000)
        Dn, In = idx_nccl.search(xq, 1000)
        assert np.array_equal(I, In) and np.array_equal(D, Dn)
        if idx._peer is not None:
            idx._peer.check()
        # collective .c2df ingest (config C5's loader) + SGI2 shard files: save, reload, same answers
        from oracle import c2df_ref
        paths = sorted(str(p) for p in Path(out_dir).glob("corpus/*.c2df"))
        idx2 = ShardedIndexFlatIP(d)
        status = idx2.add_c2df_paths(paths, n_threads=2)
        good = [i for i, p in enumerate(paths) if "bad" not in Path(p).name]
        assert list(np.nonzero(status == 0)[0]) == good and idx2.ntotal == len(good)
        rows = np.stack([c2df_ref.decode_clip(Path(paths[i]).read_bytes())[1] for i in good])
        qs = rows[[0, len(good) // 2, len(good) - 1]]
        D2, I2 = idx2.search(qs, 5)
        assert list(I2[:, 0]) == [0, len(good) // 2, len(good) - 1] and np.all(np.abs(D2[:, 0] - 1.0) < 2e-3)
        check_topk(D2, I2, rows.astype(np.float16).astype(np.float64), qs.astype(np.float16).astype(np.float64), 5,
                   score_tol=1e-3)
        idx2.save(Path(out_dir) / "shards")
        idx3 = ShardedIndexFlatIP.load(Path(out_dir) / "shards")
        assert idx3.ntotal == idx2.ntotal and idx3.local_ntotal == idx2.local_ntotal
        D3, I3 = idx3.search(qs, 5)
        assert np.array_equal(I3, I2) and np.array_equal(D3, D2)
        (Path(out_dir) / f"ok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


def _write_c2df_corpus(root, d=512, n=301):
    from sgic_b200 import c2df
    from sgic_b200.index_build import quantize_u8_and_compress
    rng = np.random.default_rng(5)
    (root / "corpus").mkdir()
    for i in range(n):
        v = rng.standard_normal(d).astype(np.float32)
        v /= np.linalg.norm(v)
        payload, meta = quantize_u8_and_compress(v)
        name = f"img_{i:04d}.c2df"
        blob = c2df.pack_c2df({"clip_stream": payload, "clip_meta": meta}, {"version": 2})
        if i in (2, 140, 170, 171):                            # broken files in both ranks' slices
            name, blob = f"img_{i:04d}_bad.c2df", b"XXXX" + blob[4:]
        (root / "corpus" / name).write_bytes(blob)


def test_two_gpu_nccl_sharded_search(tmp_path):
    _write_c2df_corpus(tmp_path)
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
    # one box, NVLink between the GPUs: the peer exchange must have come up (a silent NCCL-only run would hide it)
    assert (tmp_path / "exchange0").read_text() == "peer" and (tmp_path / "exchange1").read_text() == "peer"
