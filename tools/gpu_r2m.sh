#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r03f}
timeout 400 python -m pytest tests/test_gpu_ppo_pipeline.py -m gpu -x -q 2>&1 | tail -3
for ARGS in "--attn auto" "--attn auto --layout bloxCapture" "--attn auto --layout tinyCapture"; do
  timeout 600 python tools/bench_mappo.py --steps 32 --updates 3 $ARGS 2>&1 | tail -1 | cut -c1-330 | tee -a gpurun_out/mappo_$TAG.log
done
