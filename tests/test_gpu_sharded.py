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
        for nq, k in ((1, 10), (6, 10), (2, 100)):
            xq = np.concatenate([xb[:1], unit(rng, nq - 1, d)]) if nq > 1 else xb[:1].copy()
            D, I = idx.search(xq, k)
            xb16 = xb.astype(np.float16).astype(np.float64)
            check_topk(D, I, xb16, xq.astype(np.float16).astype(np.float64), k, score_tol=2e-5, tie_tol=1e-6)
            assert I[0, 0] == 0 and I[0, 1] == 40_000      # exact tie -> lower global id first
        (Path(out_dir) / f"ok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


def test_two_gpu_nccl_sharded_search(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
