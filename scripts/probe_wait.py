"""Does a sleeping wait beat a polling one on a power-capped part?  100M x 512, alternating settings on ONE index:
K3 (batch 1) with consumers sleeping 0 / 20 / 50 / 100 ns after a failed mbarrier try, and the batch-4096 kernel with
epilogue warps sleeping 0 / 50 / 200 ns.  3-second loops, nvidia-smi sampled.  usage: probe_wait.py [rows]"""
import statistics, subprocess, sys, threading, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from sgic_b200 import faiss_compat as faiss
from sgic_b200.synth import fill_index_random, random_unit_queries

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
d, k = 512, 10
idx = faiss.IndexFlatIP(d, device=0, retain_fp32=False)
fill_index_random(idx, n, chunk_rows=500_000)
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                       stdout=subprocess.PIPE, text=True)
samples = []
def reader():
    for line in smi.stdout:
        try:
            a, b = line.strip().split(",")
            samples.append((time.time(), float(a), float(b)))
        except Exception:
            pass
threading.Thread(target=reader, daemon=True).start()
def run(nq, secs, name):
    q = torch.from_numpy(random_unit_queries(nq, d)).cuda()
    D = torch.empty((nq, k), device="cuda"); I = torch.empty((nq, k), dtype=torch.int64, device="cuda")
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
    clk = [s for (t, s, p) in samples if t0 + 0.5 < t < t1]; pw = [p for (t, s, p) in samples if t0 + 0.5 < t < t1]
    print(f"{name:34s} ms={ms:9.3f} GB/s={n*d*2/ms/1e6:6.0f} TF={2*nq*n*d/ms/1e9:6.0f} sm_mhz={statistics.median(clk) if clk else None} power={statistics.median(pw) if pw else None}", flush=True)
print(f"rows={n}", flush=True)
for rep in range(2):
    for ns in (0, 20, 50, 100, 200):
        idx.set_option("scan_wait_ns", ns)
        run(1, 3.0, f"K3 batch 1, consumers sleep {ns} ns")
idx.set_option("scan_wait_ns", 0)
for rep in range(2):
    for ns in (0, 50, 200, 1000):
        idx.set_option("epi_wait_ns", ns)
        run(4096, 3.0, f"batch 4096, epilogue sleeps {ns} ns")
idx.set_option("epi_wait_ns", 0)
smi.kill()
