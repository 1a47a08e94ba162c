#!/bin/bash
for f in "" "-DPM_TMA_STORE=1"; do
  echo "== variant: $f"
  PM_NVCC_EXTRA="$f" python -m pacman_marl_2025_b200.build --force > /dev/null || exit 1
  for l in tinyCapture smallCapture bloxCapture; do python tools/time_rules_only.py $l 2>&1 | grep "with obs"; done
  python tools/time_rules_only.py tinyCapture 4096 2>&1 | grep "with obs"
  python tools/time_observe.py
done
echo "== parity tests on the TMA build"
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fixture or large_batch or auto_reset or max_size or per_env or properties" 2>&1 | tail -3
PM_NVCC_EXTRA="" python -m pacman_marl_2025_b200.build --force > /dev/null
