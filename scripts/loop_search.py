"""Runs one batch size in a loop for a few seconds (for clock/power sampling). usage: loop_search.py rows nq seconds [dense_min_nq]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from sgic_b200 import faiss_compat as faiss
from sgic_b200.synth import fill_index_random, random_unit_queries
n, nq, secs = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
idx = faiss.IndexFlatIP(512, device=0, retain_fp32=False); fill_index_random(idx, n)
if len(sys.argv) > 4: idx.set_option("dense_min_nq", int(sys.argv[4]))
q = torch.from_numpy(random_unit_queries(nq, 512)).cuda()
D = torch.empty((nq, 10), device="cuda"); I = torch.empty((nq, 10), dtype=torch.int64, device="cuda")
time.sleep(1.0)
t0 = time.time(); it = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < secs:
    for _ in range(10): idx.search_torch(q, 10, out=(D, I))
    torch.cuda.synchronize(); it += 10
e1.record(); torch.cuda.synchronize()
print(f"nq={nq} rows={n}: {e0.elapsed_time(e1)/it:.3f} ms/search over {it} searches", flush=True)
