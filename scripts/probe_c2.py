"""Breakdown of the C2 (1Mx512, batch 1) search: scan kernel vs whole call, CUDA events."""
import sys, statistics
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from sgic_b200 import faiss_compat as faiss
from sgic_b200.synth import fill_index_random, random_unit_queries
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
idx = faiss.IndexFlatIP(512, device=0, retain_fp32=False)
fill_index_random(idx, n)
q = torch.from_numpy(random_unit_queries(1, 512)).cuda()
D = torch.empty((1, 10), device="cuda"); I = torch.empty((1, 10), dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for fused in (1, 0):
    idx.set_option("fused", fused)
    idx.set_option("timing", 1)
    for _ in range(5): idx.search_torch(q, 10, out=(D, I))
    sc, se = [], []
    for _ in range(30):
        idx.search_torch(q, 10, out=(D, I)); sc.append(idx.stat("last_scan_ns")); se.append(idx.stat("last_search_ns"))
    idx.set_option("timing", 0)
    # back-to-back untimed-inside loop (what bench `value` sees) and L2-flushed single calls
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): idx.search_torch(q, 10, out=(D, I))
    e1.record(); torch.cuda.synchronize()
    b2b = e0.elapsed_time(e1) / 200 * 1e3
    fl = []
    for _ in range(30):
        flush.zero_()
        e0.record(); idx.search_torch(q, 10, out=(D, I)); e1.record(); torch.cuda.synchronize()
        fl.append(e0.elapsed_time(e1) * 1e3)
    print(f"fused={fused} scan_us={statistics.median(sc)/1e3:.1f} search_us={statistics.median(se)/1e3:.1f} "
          f"b2b_us={b2b:.1f} flushed_us={statistics.median(fl):.1f} -> {n*1024/statistics.median(fl)/1e3:.0f} GB/s "
          f"({n*1024/statistics.median(fl)/1e3/6500.6:.3f} of measured)")
