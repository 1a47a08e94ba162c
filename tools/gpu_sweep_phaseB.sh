#!/bin/bash
# build + time the step kernel for each phase-B variant (dev sweep)
mkdir -p gpurun_out
./tools/write_peak 2>&1 | grep -E "256-bit|grid 2048" 
for f in "" "-DPM_ST256=1" "-DPM_ALU_EXPAND=1" "-DPM_ST256=1 -DPM_ALU_EXPAND=1"; do
  echo "== variant: $f"
  PM_NVCC_EXTRA="$f" python -m pacman_marl_2025_b200.build --force > /dev/null || exit 1
  for l in tinyCapture smallCapture bloxCapture; do python tools/time_rules_only.py $l 2>&1 | grep "with obs"; done
done
PM_NVCC_EXTRA="" python -m pacman_marl_2025_b200.build --force > /dev/null
