"""Summarise an .ncu-rep: headline metrics + per-SASS-instruction stall samples (top N, and per window)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw))); h = r[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_uniform.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "lts__t_sector_hit_rate.pct"]
for i, c in enumerate(h):
    if c in want or ("tensor" in c and "pct_of_peak_sustained_elapsed" in c and ".avg." in c):
        print(c, [row[i][:60] for row in r[1:]])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(src)))
hi = next(i for i, row in enumerate(r) if "Source" in row and "# Samples" in row)
h = r[hi]; rows = r[hi + 1:]
iS = h.index("# Samples"); iSrc = h.index("Source"); iEx = h.index("Instructions Executed")
stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(x[iS]) for x in rows)
print("total samples", tot, "instructions", len(rows))
top = sorted(range(len(rows)), key=lambda i: -int(rows[i][iS]))[:topn]
for i in sorted(top):
    x = rows[i]
    st = sorted(((h[j][6:], int(x[j])) for j in stall if int(x[j]) > 0), key=lambda kv: -kv[1])[:3]
    print(i, x[iS], x[iEx], x[iSrc].strip()[:80], st)
print("-- windows of 50")
for a in range(0, len(rows), 50):
    s = sum(int(x[iS]) for x in rows[a:a + 50]); e = sum(int(x[iEx]) for x in rows[a:a + 50])
    if s: print(a, s, e, rows[a][iSrc].strip()[:50])
