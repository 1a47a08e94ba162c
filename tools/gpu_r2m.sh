#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r03c}
timeout 400 python -m pytest tests/test_gpu_ppo_pipeline.py -m gpu -x -q 2>&1 | tail -8
