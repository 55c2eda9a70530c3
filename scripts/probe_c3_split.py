"""C3 (10M x 768 fp16, batch 4096, k = 100) under sustained load: how much of a search is the dense scan kernel and
how much is everything around it (sample pass, query conversion, merge of the per-slice lists).
usage: probe_c3_split.py [rows] [seconds]"""
import statistics, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from sgic_b200 import faiss_compat as faiss
from sgic_b200.synth import fill_index_random, random_unit_queries

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
d, nq = 768, 4096
idx = faiss.IndexFlatIP(d, device=0, retain_fp32=False)
fill_index_random(idx, n)
q = torch.from_numpy(random_unit_queries(nq, d)).cuda()
for k, seed in ((100, 1), (100, 0), (10, 1)):
    idx.set_option("dense_seed", seed)
    D = torch.empty((nq, k), device="cuda"); I = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    for _ in range(3): idx.search_torch(q, k, out=(D, I))
    torch.cuda.synchronize()
    l0 = idx.stat("launches")
    idx.set_option("timing", 2)
    idx.scan_times_ms()
    t0 = time.time(); it = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(5): idx.search_torch(q, k, out=(D, I))
        torch.cuda.synchronize(); it += 5
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / it
    scan = idx.scan_times_ms(min(it, 64))
    idx.set_option("timing", 0)
    print(f"k={k} seed={seed}: search {ms:.3f} ms, scan kernel {statistics.mean(scan):.3f} ms (n={len(scan)}), "
          f"around it {ms - statistics.mean(scan):.3f} ms, launches per search {(idx.stat('launches') - l0) / it:.1f}", flush=True)
