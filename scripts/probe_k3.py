"""GPU probe: K3 batch-small throughput vs HBM roofline at several sizes / options."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from sgic_b200 import faiss_compat as faiss
from sgic_b200.synth import fill_index_random, random_unit_queries

def timeit(fn, iters, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def main():
    import os
    quick = os.environ.get("PROBE_QUICK") == "1"
    sizes = [int(a) for a in sys.argv[1:]] or [1_000_000, 16_000_000]
    d, k = 512, 10
    out = []
    for n in sizes:
        idx = faiss.IndexFlatIP(d, device=0, retain_fp32=False)
        t0 = time.time(); fill_index_random(idx, n); t_fill = time.time() - t0
        for nq in ((1,) if quick else (1, 2, 4)):
            q = torch.from_numpy(random_unit_queries(nq, d)).cuda()
            D = torch.empty((nq, k), dtype=torch.float32, device="cuda"); I = torch.empty((nq, k), dtype=torch.int64, device="cuda")
            for opts in ({}, {"stages": 3}, {"stages": 4}, {"stages": 5}, {"stages": 6}, {"rb": 8, "stages": 3}, {"rb": 8, "stages": 2}, {"rb": 2, "stages": 8}, {"rb": 2, "stages": 6}, {"rb": 2, "stages": 4}, {"grid": 296, "stages": 4}, {"grid": 444, "stages": 4}):
                if (nq > 1 or quick) and opts: continue
                for kk, v in (("stages", 0), ("grid", 0), ("evict_first", 1), ("rb", 0)): idx.set_option(kk, v)
                for kk, v in opts.items(): idx.set_option(kk, v)
                try:
                    ms = timeit(lambda: idx.search_torch(q, k, out=(D, I)), 5 if quick else (20 if n > 4_000_000 else 200))
                except Exception as e:
                    print("ERR", n, nq, opts, e); continue
                gbs = n * d * 2 / ms / 1e6
                rec = dict(n=n, nq=nq, opts=opts, ms=round(ms, 4), GBs=round(gbs, 1), frac_meas=round(gbs / 6500.6, 3),
                           qps=round(nq / ms * 1e3, 1), grid=idx.stat("last_grid"), stages=idx.stat("last_stages"))
                print(json.dumps(rec), flush=True); out.append(rec)
        # host-API end-to-end
        qh = random_unit_queries(1, d)
        for _ in range(3): idx.search(qh, k)
        t0 = time.perf_counter(); it = 50
        for _ in range(it): idx.search(qh, k)
        e2e = (time.perf_counter() - t0) / it * 1e3
        print(json.dumps(dict(n=n, e2e_ms=round(e2e, 4), fill_s=round(t_fill, 2))), flush=True)
        idx.close(); torch.cuda.empty_cache()
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/probe_k3.json").write_text(json.dumps(out, indent=1))

if __name__ == "__main__":
    main()
