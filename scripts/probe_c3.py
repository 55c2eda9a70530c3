"""C3 (10M x 768 fp16, batch 4096) taken apart under sustained load: k = 100 as configured, the same with the
epilogue switched off (debug = 4: wrong answers, timing only), and k = 10 (thread-private lists instead of
reservoirs) — ms per search, TFLOP/s, SM clock and power (nvidia-smi) for each, 3-second loops.
usage: probe_c3.py [rows] [seconds]"""
import statistics, subprocess, sys, threading, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from sgic_b200 import faiss_compat as faiss
from sgic_b200.synth import fill_index_random, random_unit_queries

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
d, nq = 768, 4096
idx = faiss.IndexFlatIP(d, device=0, retain_fp32=False)
fill_index_random(idx, n)
q = torch.from_numpy(random_unit_queries(nq, d)).cuda()
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
print(f"rows={n} d={d} nq={nq}", flush=True)
for k, dbg, seed, slices, name in ((100, 0, 1, 0, "k=100: sample pass + early compaction"), (100, 8, 1, 0, "k=100: sample pass only"),
                                  (100, 0, 0, 0, "k=100: early compaction only"), (100, 8, 0, 0, "k=100: neither (round 1)"),
                                  (100, 4, 0, 0, "k=100, epilogue off"), (10, 0, 1, 0, "k=10"),
                                  (100, 0, 1, 0, "k=100: both, again"), (100, 8, 0, 0, "k=100: neither, again")):
    idx.set_option("dense_slices", slices)
    D = torch.empty((nq, k), device="cuda"); I = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    idx.set_option("debug", dbg)
    idx.set_option("dense_seed", seed)
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
    print(f"{name:40s} ms={ms:8.3f} TF={2*nq*n*d/ms/1e9:6.0f} sm_mhz={statistics.median(clk) if clk else None} "
          f"power={statistics.median(pw) if pw else None}", flush=True)
idx.set_option("debug", 0)
smi.kill()
