"""K4b query-ring depth A/B on ONE 100M x 512 index: "stages" = 4 (round 1) against the deepest ring that fits (5),
batch 4096 and 1024, alternating, 3-second loops with nvidia-smi sampling.  usage: probe_k4b_stages.py [rows]"""
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
print(f"rows={n}", flush=True)
ref = {}
for rep in range(2):
    for nq in (4096, 1024):
        q = torch.from_numpy(random_unit_queries(nq, d)).cuda()
        D = torch.empty((nq, k), device="cuda"); I = torch.empty((nq, k), dtype=torch.int64, device="cuda")
        for stages in (4, 0):
            idx.set_option("stages", stages)
            for _ in range(2): idx.search_torch(q, k, out=(D, I))
            torch.cuda.synchronize()
            if nq in ref: assert torch.equal(ref[nq], I), "answers differ between ring depths"
            ref[nq] = I.clone()
            t0 = time.time(); it = 0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            while time.time() - t0 < 3.0:
                for _ in range(3): idx.search_torch(q, k, out=(D, I))
                torch.cuda.synchronize(); it += 3
            e1.record(); torch.cuda.synchronize()
            t1 = time.time()
            ms = e0.elapsed_time(e1) / it
            clk = [s for (t, s, p) in samples if t0 + 0.5 < t < t1]; pw = [p for (t, s, p) in samples if t0 + 0.5 < t < t1]
            print(f"nq={nq:5d} ring={idx.stat('last_stages')} kernel={idx.stat('last_kernel')} ms={ms:9.3f} TF={2*nq*n*d/ms/1e9:6.0f} "
                  f"sm_mhz={statistics.median(clk) if clk else None} power={statistics.median(pw) if pw else None}", flush=True)
idx.set_option("stages", 0)
smi.kill()
