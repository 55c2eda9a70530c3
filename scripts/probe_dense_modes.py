"""Timing of the dense path (K4) per batch size / kernel variant, CUDA events, no profiler.
usage: probe_dense_modes.py [n] [d] [k] [nq,nq,...] [mode,mode,...] [dbg,dbg,...]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from sgic_b200 import faiss_compat as faiss
from sgic_b200.synth import fill_index_random, random_unit_queries

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
nqs = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1, 4, 8, 16, 64, 128, 256, 1024, 4096]
modes = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [0]
dbgs = [int(x) for x in sys.argv[6].split(",")] if len(sys.argv) > 6 else [0, 4]
l2s = [int(x) for x in sys.argv[7].split(",")] if len(sys.argv) > 7 else [0]
if len(sys.argv) > 8:
    min_nq = int(sys.argv[8])
else:
    min_nq = None
names = {0: "auto", 1: "1cta", 2: "pairs+stream", 3: "pairs", 4: "pairs+dbres"}
idx = faiss.IndexFlatIP(d, device=0, retain_fp32=False)
fill_index_random(idx, n)
if min_nq is not None:
    idx.set_option("dense_min_nq", min_nq)
if len(sys.argv) > 9:
    idx.set_option("stages", int(sys.argv[9]))

import os
for kv in filter(None, os.environ.get("SGIC_PROBE_OPTS", "").split(",")):   # e.g. SGIC_PROBE_OPTS=dense_gthr=0
    name, val = kv.split("=")
    idx.set_option(name, int(val))
print(f"n={n} d={d} k={k} dense_min_nq={min_nq} opts={os.environ.get('SGIC_PROBE_OPTS', '')}", flush=True)
for nq in nqs:
    q = torch.from_numpy(random_unit_queries(nq, d)).cuda()
    D = torch.empty((nq, k), device="cuda")
    I = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    for mode in modes:
      for l2 in l2s:
        for dbg in dbgs:
            if l2:
                idx.set_option("dense_l2_mb", l2)
            idx.set_option("dense_mode", mode)
            idx.set_option("debug", dbg)
            for _ in range(2):
                idx.search_torch(q, k, out=(D, I))
            torch.cuda.synchronize()
            reps = 3 if nq >= 1024 else 10
            samples = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    idx.search_torch(q, k, out=(D, I))
                e1.record()
                torch.cuda.synchronize()
                samples.append(e0.elapsed_time(e1) / reps)
            samples.sort()
            ms, best = samples[len(samples) // 2], samples[0]
            idx.set_option("timing", 1)
            sc, tot = [], []
            for _ in range(5):
                idx.search_torch(q, k, out=(D, I))
                sc.append(idx.stat("last_scan_ns") / 1e6)
                tot.append(idx.stat("last_search_ns") / 1e6)
            idx.set_option("timing", 0)
            tf = 2 * nq * n * d / ms / 1e9
            print(f"nq={nq:5d} mode={names[mode]:12s} l2={l2:3d} skipEpi={dbg >> 2} ms={ms:8.3f} (best {best:8.3f}) TF={tf:7.0f} "
                  f"({tf / 1404.9:.3f} of sustained)  GB/s={n * d * 2 / ms / 1e6:6.0f}  qps={nq / ms * 1e3:9.0f}  [timed alone: scan {min(sc):.3f} total {min(tot):.3f}]", flush=True)
