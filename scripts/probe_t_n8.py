"""K4t with MMA N = 8 against N = 16 for batches of up to 8 queries (option "t_n8"), same index, alternating."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from sgic_b200 import faiss_compat as faiss
from sgic_b200.synth import fill_index_random, random_unit_queries
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
idx = faiss.IndexFlatIP(512, device=0, retain_fp32=False)
fill_index_random(idx, n)
idx.set_option("dense_min_nq", 1)
for nq in (1, 2, 8):
    q = torch.from_numpy(random_unit_queries(nq, 512)).cuda()
    D = torch.empty((nq, 10), device="cuda"); I = torch.empty((nq, 10), dtype=torch.int64, device="cuda")
    ref = None
    for n8 in (0, 1, 0, 1):
        idx.set_option("t_n8", n8)
        for _ in range(30): idx.search_torch(q, 10, out=(D, I))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(40): idx.search_torch(q, 10, out=(D, I))
        e1.record(); torch.cuda.synchronize()
        if ref is None: ref = (D.clone(), I.clone())
        same = bool(torch.equal(I, ref[1]) and torch.allclose(D, ref[0], atol=1e-6))
        print(f"nq={nq} t_n8={n8}: {e0.elapsed_time(e1) / 40:.3f} ms  kernel={idx.stat('last_kernel')} same_as_first={same}", flush=True)
