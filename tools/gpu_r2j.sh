#!/bin/bash
# ncu (full set + source) of the plain and bots step kernels; reports kept for tools/ncu_lines.py
TAG=${1:-r02j}
mkdir -p gpurun_out
for V in plain bots; do
  CMD="python tools/run_variants.py --reps 1 --warm 2 --small 0 --only $V"
  $CMD > gpurun_out/plain_$V.log 2>&1 &&
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:pm_env_kernelILi0E -s 2 -c 1 -f \
      -o gpurun_out/prof_${V}_$TAG $CMD > gpurun_out/ncu_${V}_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_${V}_$TAG.log
  ncu -i gpurun_out/prof_${V}_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_raw_${V}_$TAG.csv 2>/dev/null
done
ls -la gpurun_out/*$TAG*
