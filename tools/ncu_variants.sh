#!/bin/bash
# ncu of the step-kernel variants: one launch each, full set + source; the reports are converted to CSV pages on the
# box and deleted (a report with source is ~15 MB per launch; gpurun_out/ is capped at 64 MiB).
TAG=${1:-r02}
mkdir -p gpurun_out
for V in plain views views+obs bots bots+views; do
  CMD="python tools/run_variants.py --reps 1 --warm 2 --small 0 --only $V"
  $CMD > gpurun_out/plain_$V.log 2>&1 &&
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:pm_env_kernelILi0E -s 2 -c 1 -f \
      -o gpurun_out/prof_${V}_$TAG $CMD > gpurun_out/ncu_${V}_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_${V}_$TAG.log
  if [ -f gpurun_out/prof_${V}_$TAG.ncu-rep ]; then
    ncu -i gpurun_out/prof_${V}_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_raw_${V}_$TAG.csv 2>/dev/null
    ncu -i gpurun_out/prof_${V}_$TAG.ncu-rep --page details > gpurun_out/ncu_details_${V}_$TAG.txt 2>/dev/null
    true || rm -f gpurun_out/prof_${V}_$TAG.ncu-rep
  fi
done
ls -la gpurun_out | tail -30; du -sh gpurun_out
