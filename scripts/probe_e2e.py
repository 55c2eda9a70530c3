"""Where the end-to-end batch-1 call spends its time: host API (numpy in/out) vs device-resident call + sync.
usage: probe_e2e.py [rows]"""
import sys, time, statistics
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from sgic_b200 import faiss_compat as faiss
from sgic_b200.synth import fill_index_random, random_unit_queries
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
idx = faiss.IndexFlatIP(512, device=0, retain_fp32=False)
fill_index_random(idx, n)
qh = random_unit_queries(1, 512)
q = torch.from_numpy(qh).cuda()
D = torch.empty((1, 10), device="cuda"); I = torch.empty((1, 10), dtype=torch.int64, device="cuda")
def wall(fn, reps=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(ts), min(ts), max(ts)
def b2b():
    for _ in range(20): idx.search_torch(q, 10, out=(D, I))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): idx.search_torch(q, 10, out=(D, I))
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20
def dev_sync():
    idx.search_torch(q, 10, out=(D, I)); torch.cuda.synchronize()
for rnd in range(2):
    print(f"back-to-back device loop      : {b2b():.3f} ms/step")
    print("device call + synchronize     : median %.3f min %.3f max %.3f ms" % wall(dev_sync))
    print("host API index.search(numpy)  : median %.3f min %.3f max %.3f ms" % wall(lambda: idx.search(qh, 10)))
    idx.set_option("timing", 1)
    idx.search(qh, 10)
    print(f"  kernel inside a host call   : scan {idx.stat('last_scan_ns')/1e6:.3f} ms, search {idx.stat('last_search_ns')/1e6:.3f} ms")
    idx.set_option("timing", 0)
