#!/bin/bash
PM_NVCC_EXTRA="-DPM_TRACE" python -m pacman_marl_2025_b200.build --force > /dev/null || exit 1
python tools/trace_step.py tinyCapture
PM_NVCC_EXTRA="" python -m pacman_marl_2025_b200.build --force > /dev/null
