#!/bin/bash
# usage: probe_clocks.sh <label> <command...>  -- runs the command while sampling SM clock / power every 100 ms
label=$1; shift
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader,nounits -lms 100 > gpurun_out/clk_$label.csv &
SMI=$!
"$@"
kill $SMI
python - <<PY
import statistics
rows=[l.strip().split(',') for l in open('gpurun_out/clk_$label.csv') if l.strip()]
sm=[float(r[0]) for r in rows]; pw=[float(r[2]) for r in rows]
busy=[(s,p) for s,p in zip(sm,pw) if p>400]
print('$label', 'samples',len(rows),'busy',len(busy),'sm_mhz median(busy)', statistics.median([s for s,_ in busy]) if busy else None, 'power median(busy)', statistics.median([p for _,p in busy]) if busy else None, 'cap active', sum('Active' in r[3] and 'Not' not in r[3] for r in rows))
PY
