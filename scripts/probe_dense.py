"""GPU probe: K4 dense path throughput (TFLOP/s vs the measured bf16 dense peak) and HBM rate."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from sgic_b200 import faiss_compat as faiss
from sgic_b200.synth import fill_index_random, random_unit_queries

def timeit(fn, iters, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

cfgs = [(10_000_000, 512, 10, [5, 8, 16, 64, 128, 256, 1024, 4096]), (10_000_000, 768, 100, [4096])]
if len(sys.argv) > 1 and sys.argv[1] == "big":
    cfgs = [(100_000_000, 512, 10, [8, 128, 1024, 4096])]
for n, d, k, nqs in cfgs:
    idx = faiss.IndexFlatIP(d, device=0, retain_fp32=False)
    fill_index_random(idx, n)
    for nq in nqs:
        q = torch.from_numpy(random_unit_queries(nq, d)).cuda()
        D = torch.empty((nq, k), dtype=torch.float32, device="cuda"); I = torch.empty((nq, k), dtype=torch.int64, device="cuda")
        iters = 3 if nq * n > 2e10 else 10
        ms = timeit(lambda: idx.search_torch(q, k, out=(D, I)), iters)
        tf = 2.0 * nq * n * d / ms / 1e9
        gbs = n * d * 2 / ms / 1e6
        print(json.dumps(dict(n=n, d=d, k=k, nq=nq, ms=round(ms, 3), qps=round(nq / ms * 1e3, 1), TFLOPs=round(tf, 1),
                              frac_tc_sustained=round(tf / 1404.9, 3), GBs=round(gbs, 1), frac_hbm=round(gbs / 6500.6, 3),
                              grid=idx.stat("last_grid"))), flush=True)
    idx.close(); torch.cuda.empty_cache()
