"""Same-box A/B of two builds of libsgic.so (SGIC_LIB): each build runs in its own process on a fresh 100M x 512 index,
batch 4096 / 1024 / 1, 3-second loops; the builds alternate twice.  usage: probe_libs.py libA.so libB.so"""
import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
CHILD = r'''
import sys, time, statistics
sys.path.insert(0, %r)
import os, ctypes, torch
from sgic_b200 import _native
_L = ctypes.CDLL(os.environ["SGIC_LIB"])   # an older build exports fewer entry points: bind what is there
_native._SIGNATURES = {k: v for k, v in _native._SIGNATURES.items() if hasattr(_L, k)}
from sgic_b200 import faiss_compat as faiss
from sgic_b200.synth import fill_index_random, random_unit_queries
n, d, k = 100_000_000, 512, 10
idx = faiss.IndexFlatIP(d, device=0, retain_fp32=False)
fill_index_random(idx, n, chunk_rows=500_000)
for nq in (4096, 1024, 1):
    q = torch.from_numpy(random_unit_queries(nq, d)).cuda()
    D = torch.empty((nq, k), device="cuda"); I = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    for _ in range(3): idx.search_torch(q, k, out=(D, I))
    torch.cuda.synchronize()
    t0 = time.time(); it = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(3): idx.search_torch(q, k, out=(D, I))
        torch.cuda.synchronize(); it += 3
    e1.record(); torch.cuda.synchronize()
    print(f"  nq={nq:5d} kernel={idx.stat('last_kernel')} ring={idx.stat('last_stages')} ms={e0.elapsed_time(e1)/it:9.3f}", flush=True)
idx.close()
n, d, k, nq = 10_000_000, 768, 100, 4096        # C3
idx = faiss.IndexFlatIP(d, device=0, retain_fp32=False)
fill_index_random(idx, n)
q = torch.from_numpy(random_unit_queries(nq, d)).cuda()
D = torch.empty((nq, k), device="cuda"); I = torch.empty((nq, k), dtype=torch.int64, device="cuda")
for _ in range(3): idx.search_torch(q, k, out=(D, I))
torch.cuda.synchronize()
t0 = time.time(); it = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 3.0:
    for _ in range(5): idx.search_torch(q, k, out=(D, I))
    torch.cuda.synchronize(); it += 5
e1.record(); torch.cuda.synchronize()
print(f"  C3 10Mx768 nq=4096 k=100 ms={e0.elapsed_time(e1)/it:9.3f}", flush=True)
''' % str(ROOT)
libs = sys.argv[1:3]
for rep in range(2):
    for lib in libs:
        print(f"== {lib}", flush=True)
        subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, SGIC_LIB=str(Path(lib).resolve())), check=False)
