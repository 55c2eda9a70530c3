import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from sgic_b200 import faiss_compat as faiss
from sgic_b200.synth import fill_index_random, random_unit_queries
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
d = int(sys.argv[3]) if len(sys.argv) > 3 else 512
k = int(sys.argv[4]) if len(sys.argv) > 4 else 10
idx = faiss.IndexFlatIP(d, device=0, retain_fp32=False); fill_index_random(idx, n)
if len(sys.argv) > 5:
    idx.set_option("dense_l2_mb", int(sys.argv[5]))
if len(sys.argv) > 6:
    idx.set_option("dense_mode", int(sys.argv[6]))

q = torch.from_numpy(random_unit_queries(nq, d)).cuda()
for _ in range(3):
    D, I = idx.search_torch(q, k)
torch.cuda.synchronize()
print("ok", float(D[0, 0]))
