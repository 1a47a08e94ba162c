#!/bin/bash
for f in "" "-DPM_MIN_CTAS_BOTS=6" "-DPM_MIN_CTAS_BOTS=8"; do
  echo "== variant: $f"
  PM_NVCC_EXTRA="$f -Xptxas -v" python -m pacman_marl_2025_b200.build --force 2>&1 | grep -A2 "pm_env_kernelILi0ELi1ELi0" | grep -E "spill|Used" | tr '\n' ' '; echo
  python tools/time_bots.py
done
PM_NVCC_EXTRA="" python -m pacman_marl_2025_b200.build --force > /dev/null
