#!/bin/bash
PM_NVCC_EXTRA="-DPM_TRACE" python -m pacman_marl_2025_b200.build --force > /dev/null || exit 1
python tools/trace_step.py ${1:-tinyCapture} | grep -v "^   cta" 
python tools/trace_step.py ${1:-tinyCapture} | grep -A16 "^SM 0:"
PM_NVCC_EXTRA="" python -m pacman_marl_2025_b200.build --force > /dev/null
