#!/bin/bash
for f in "" "-DPM_MIN_CTAS_OBS=12" "-DPM_MIN_CTAS_OBS=16" "-DPM_MIN_CTAS_OBS=16 -DPM_ST256=1" "-DPM_MIN_CTAS_OBS=12 -DPM_ST256=1"; do
  echo "== variant: $f"
  PM_NVCC_EXTRA="$f -Xptxas -v" python -m pacman_marl_2025_b200.build --force 2>&1 | grep -A2 "pm_env_kernelILi2ELi0ELi0" | grep -E "spill|Used" | tr '\n' ' '; echo
  python tools/time_observe.py
done
PM_NVCC_EXTRA="" python -m pacman_marl_2025_b200.build --force > /dev/null
