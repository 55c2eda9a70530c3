import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from sgic_b200 import faiss_compat as faiss
from sgic_b200.synth import fill_index_random, random_unit_queries
n, d, k = 10_000_000, 512, 10
idx = faiss.IndexFlatIP(d, device=0, retain_fp32=False); fill_index_random(idx, n)
for nq in (4096, 128):
    q = torch.from_numpy(random_unit_queries(nq, d)).cuda()
    D = torch.empty((nq, k), device="cuda"); I = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    for dbg in (0, 4, 1, 2, 3, 5, 6, 7):
        idx.set_option("debug", dbg)
        for _ in range(2): idx.search_torch(q, k, out=(D, I))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): idx.search_torch(q, k, out=(D, I))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"nq={nq} debug={dbg} (skipA={dbg&1} skipB={(dbg>>1)&1} skipEpi={(dbg>>2)&1}) ms={ms:.2f} TF={2*nq*n*d/ms/1e9:.0f}", flush=True)
