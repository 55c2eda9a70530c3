"""Mid-size batches on a large shard: ms per search, GB/s and the SM clock / power under load per batch size.
usage: probe_mid.py rows nq,nq,... [dbg,dbg] [mode,mode] [seconds] [k] [dense_min_nq,...]"""
import subprocess, sys, time, statistics
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from sgic_b200 import faiss_compat as faiss
from sgic_b200.synth import fill_index_random, random_unit_queries

n = int(sys.argv[1])
nqs = [int(x) for x in sys.argv[2].split(",")]
dbgs = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
modes = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]
secs = float(sys.argv[5]) if len(sys.argv) > 5 else 2.0
k = int(sys.argv[6]) if len(sys.argv) > 6 else 10
min_nqs = [int(x) for x in sys.argv[7].split(",")] if len(sys.argv) > 7 else [0]
d = 512
idx = faiss.IndexFlatIP(d, device=0, retain_fp32=False)
fill_index_random(idx, n)
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                       stdout=subprocess.PIPE, text=True)
import threading
samples = []
def reader():
    for line in smi.stdout:
        try:
            a, b = line.strip().split(",")
            samples.append((time.time(), float(a), float(b)))
        except Exception:
            pass
threading.Thread(target=reader, daemon=True).start()
print(f"rows={n} d={d} k={k}", flush=True)
for nq in nqs:
    q = torch.from_numpy(random_unit_queries(nq, d)).cuda()
    D = torch.empty((nq, k), device="cuda"); I = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    for mode in [(m, mn) for m in modes for mn in min_nqs]:
        mode, mn = mode
        idx.set_option("dense_min_nq", mn)
        for dbg in dbgs:
            idx.set_option("dense_mode", mode); idx.set_option("debug", dbg)
            for _ in range(3): idx.search_torch(q, k, out=(D, I))
            torch.cuda.synchronize()
            t0 = time.time(); it = 0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            while time.time() - t0 < secs:
                for _ in range(5): idx.search_torch(q, k, out=(D, I))
                torch.cuda.synchronize(); it += 5
            e1.record(); torch.cuda.synchronize()
            t1 = time.time()
            ms = e0.elapsed_time(e1) / it
            clk = [s for (t, s, p) in samples if t0 + 0.5 < t < t1]
            pw = [p for (t, s, p) in samples if t0 + 0.5 < t < t1]
            print(f"nq={nq:5d} mode={mode} min_nq={mn} dbg={dbg} ms={ms:8.3f} GB/s={n*d*2/ms/1e6:6.0f} TF={2*nq*n*d/ms/1e9:6.0f} "
                  f"sm_mhz={statistics.median(clk) if clk else None} power={statistics.median(pw) if pw else None}", flush=True)
smi.kill()
