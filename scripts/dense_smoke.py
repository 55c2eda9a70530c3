"""First-contact checks of the tcgen05 dense path (K4), smallest shapes first, each against the oracle."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from sgic_b200 import faiss_compat as faiss
from oracle.flat_ip import check_topk

def unit(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32); x /= np.linalg.norm(x, axis=1, keepdims=True); return x

cases = [(256, 64, 5, 4), (1000, 64, 5, 10), (1000, 512, 8, 10), (5000, 512, 128, 10), (5000, 512, 129, 10),
         (70000, 512, 300, 10), (20000, 768, 64, 100), (3001, 520, 17, 10), (200000, 512, 1024, 10),
         (40000, 512, 200, 32), (40000, 512, 200, 33), (9000, 256, 130, 1), (300, 512, 140, 32),
         (50000, 768, 260, 10), (100000, 128, 4096, 5), (20000, 512, 64, 1024),
         (300, 512, 140, 100), (5000, 512, 1000, 128), (150000, 768, 300, 100), (60000, 256, 129, 500)]
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
if len(sys.argv) > 2:
    cases = cases[:int(sys.argv[2])]
print("dense_mode", mode, flush=True)
for n, d, nq, k in cases:
    rng = np.random.default_rng(n + d + nq + k)
    xb, xq = unit(rng, n, d), unit(rng, nq, d)
    idx = faiss.IndexFlatIP(d, device=0)
    idx.add(xb)
    idx.set_option("dense_mode", mode)
    idx.search(xq, k)
    t0 = time.time()
    D, I = idx.search(xq, k)
    dt = time.time() - t0
    xb16 = xb.astype(np.float16).astype(np.float64); xq16 = xq.astype(np.float16).astype(np.float64)
    try:
        sel = np.arange(nq) if nq <= 64 else rng.choice(nq, 64, replace=False)
        sw = check_topk(D[sel], I[sel], xb16, xq16[sel], k, score_tol=3e-5, tie_tol=1e-6)
        print(f"OK   n={n} d={d} nq={nq} k={k}  {dt*1e3:.1f} ms  boundary swaps={sw}", flush=True)
    except AssertionError as e:
        print(f"FAIL n={n} d={d} nq={nq} k={k}: {e}", flush=True)
        s = xq16[0] @ xb16.T
        o = np.argsort(-s)[:k]
        print("  got ids", I[0][:k], "\n  ref ids", o, "\n  got D", D[0][:k], "\n  ref D", s[o])
        break
    idx.close()
