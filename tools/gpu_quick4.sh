#!/bin/bash
for l in tinyCapture smallCapture bloxCapture; do python tools/time_rules_only.py $l 2>&1 | grep "with obs"; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
PM_NVCC_EXTRA="-DPM_TRACE" python -m pacman_marl_2025_b200.build --force > /dev/null || exit 1
python tools/trace_step.py tinyCapture | grep -v "   cta"
PM_NVCC_EXTRA="" python -m pacman_marl_2025_b200.build --force > /dev/null
