"""Where the microseconds of a small K3 search go: per-CTA %globaltimer stamps (option "trace").
usage: probe_k3_trace.py [rows]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from sgic_b200 import faiss_compat as faiss
from sgic_b200.synth import fill_index_random, random_unit_queries
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
idx = faiss.IndexFlatIP(512, device=0, retain_fp32=False)
fill_index_random(idx, n)
q = torch.from_numpy(random_unit_queries(1, 512)).cuda()
D = torch.empty((1, 10), device="cuda"); I = torch.empty((1, 10), dtype=torch.int64, device="cuda")
tr = torch.zeros((148, 4), dtype=torch.int64, device="cuda")
for _ in range(50): idx.search_torch(q, 10, out=(D, I))
idx.set_option("trace", tr.data_ptr())
rows = []
for _ in range(20):
    torch.cuda.synchronize()
    idx.search_torch(q, 10, out=(D, I)); torch.cuda.synchronize()
    t = tr.cpu().numpy().astype(np.float64) / 1e3
    t0 = t[:, 0].min()
    rows.append([t[:, 0].max() - t0, np.median(t[:, 1]) - t0, t[:, 1].min() - t0, t[:, 1].max() - t0, t[:, 2].max() - t0, t[:, 3].max() - t0])
r = np.median(np.array(rows), axis=0)
print(f"rows={n}: last CTA starts +{r[0]:.1f} us | scan ends: first {r[2]:.1f} median {r[1]:.1f} last {r[3]:.1f} us | lists written {r[4]:.1f} | exit {r[5]:.1f} us"
      f"   (ideal stream {n*1024/7.4e6:.1f} us)")
idx.set_option("trace", 0)
# A/B of the shared tail ("steal") on the same index: back-to-back and one-at-a-time
for steal in (0, 1, 0, 1):
    idx.set_option("steal", steal)
    for _ in range(20): idx.search_torch(q, 10, out=(D, I))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 200 if n <= 20_000_000 else 30
    e0.record()
    for _ in range(reps): idx.search_torch(q, 10, out=(D, I))
    e1.record(); torch.cuda.synchronize()
    print(f"steal={steal}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us per search back to back", flush=True)
