#!/bin/bash
for f in "-DPM_NO_EARLY=1" "" "-DPM_EARLY_ZERO_PLANES=0"; do
  echo "== variant: $f"
  PM_NVCC_EXTRA="$f" python -m pacman_marl_2025_b200.build --force > /dev/null || exit 1
  for l in tinyCapture smallCapture bloxCapture; do python tools/time_rules_only.py $l 2>&1 | grep "with obs"; done
done
PM_NVCC_EXTRA="" python -m pacman_marl_2025_b200.build --force > /dev/null
