"""Batch-1 (K3) and small-batch (K4t) streaming rate across embedding widths and storage types.
usage: probe_dims.py [bytes_of_database]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from sgic_b200 import faiss_compat as faiss
from sgic_b200.synth import fill_index_random, random_unit_queries

target = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000_000
k = 10
for d, dtype in ((512, "fp16"), (768, "fp16"), (1024, "fp16"), (256, "fp16"), (128, "fp16"), (64, "fp16"), (384, "fp16"),
                 (1280, "fp16"), (2048, "fp16"), (512, "bf16"), (768, "bf16")):
    n = (target // (d * 2)) // 65536 * 65536
    idx = faiss.IndexFlatIP(d, dtype=dtype, device=0, retain_fp32=False)
    fill_index_random(idx, n, chunk_rows=65536)
    line = f"d={d:5d} {dtype} rows={n:10d}:"
    for nq in (1, 2, 8, 32):
        q = torch.from_numpy(random_unit_queries(nq, d)).cuda()
        D = torch.empty((nq, k), device="cuda"); I = torch.empty((nq, k), dtype=torch.int64, device="cuda")
        for _ in range(5): idx.search_torch(q, k, out=(D, I))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30): idx.search_torch(q, k, out=(D, I))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        line += f"  nq={nq}: {ms:7.3f} ms {n * d * 2 / ms / 1e6:6.0f} GB/s (kernel {idx.stat('last_kernel')})"
    print(line, flush=True)
    idx.close(); torch.cuda.empty_cache()
