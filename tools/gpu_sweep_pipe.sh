#!/bin/bash
for f in "-DPM_NO_PIPE=1" ""; do
  echo "== variant: $f"
  PM_NVCC_EXTRA="$f" python -m pacman_marl_2025_b200.build --force > /dev/null || exit 1
  for n in 1024 4096 16384 65536 262144; do python tools/time_rules_only.py tinyCapture $n 2>&1 | grep "with obs"; done
  python tools/time_rules_only.py bloxCapture 4096 2>&1 | grep "with obs"
done
PM_NVCC_EXTRA="" python -m pacman_marl_2025_b200.build --force > /dev/null
