"""Small cases of the newer kernels for compute-sanitizer (memcheck): K4b reservoirs, K4t lane lists, bounded K3 passes,
pipelined ingest.  usage: compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from sgic_b200 import faiss_compat as faiss

rng = np.random.default_rng(3)
def unit(n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)

xb = unit(3001, 512)
idx = faiss.IndexFlatIP(512, device=0)
idx.add(xb)
for mode, nq, k in ((4, 300, 10), (4, 300, 20), (4, 520, 100), (0, 300, 100), (0, 9, 10), (0, 40, 32), (0, 1, 2500), (0, 1, 10), (0, 130, 10)):
    idx.set_option("dense_mode", mode)
    D, I = idx.search(unit(nq, 512), k)
    assert I.min() >= -1 and I.max() < 3001
    print("ok", mode, nq, k, flush=True)
idx.close()
from sgic_b200 import c2df
from sgic_b200.index_build import quantize_u8_and_compress
blobs = []
for v in xb[:200]:
    payload, meta = quantize_u8_and_compress(v)
    blobs.append(c2df.pack_c2df({"clip_stream": payload, "clip_meta": meta}, {"version": 2}))
offs = np.zeros(len(blobs) + 1, dtype=np.int64)
np.cumsum([len(b) for b in blobs], out=offs[1:])
ing = faiss.IndexFlatIP(512, device=0)
added, status = ing.add_c2df(np.frombuffer(b"".join(blobs), dtype=np.uint8), offs)
assert added == 200
print("ok ingest", flush=True)
