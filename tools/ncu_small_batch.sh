#!/bin/bash
mkdir -p gpurun_out
for V in views bots+views plain; do
  CMD="python tools/run_variants.py --reps 1 --warm 2 --envs 1024 --small 0 --only $V"
  timeout 200 ncu --set full --clock-control none --kernel-name-base mangled -k regex:pm_env_kernelILi0E -s 2 -c 1 -f -o gpurun_out/prof_${V}_1024 $CMD > gpurun_out/ncu_${V}_1024.log 2>&1
  ncu -i gpurun_out/prof_${V}_1024.ncu-rep --page raw --csv > "gpurun_out/ncu_raw_${V}_tiny1024.csv" 2>/dev/null
  rm -f gpurun_out/prof_${V}_1024.ncu-rep
  python - "$V" <<'PY'
import csv, sys
v = sys.argv[1]
rows = list(csv.reader(open(f'gpurun_out/ncu_raw_{v}_tiny1024.csv')))
d = dict(zip(rows[0], rows[-1]))
print(v, d['Kernel Name'], 'us', d['gpu__time_duration.sum'], 'regs', d['launch__registers_per_thread'], 'grid', d.get('launch__grid_size'), 'occ', d['sm__warps_active.avg.pct_of_peak_sustained_active'])
PY
done
