#!/bin/bash
# Round-2 GPU call K: attention kernels - tests, critic fwd+bwd timing, MAPPO SPS
TAG=${1:-r02k}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ppo_pipeline.py -x -q 2>&1 | tail -8
timeout 600 python tools/bench_attention.py 2>&1 | tee gpurun_out/attention_$TAG.log
for ARGS in "--fused 1 --attn auto" "--fused 1 --attn auto --amp 1" "--fused 1 --attn module --amp 1"; do
  timeout 600 python tools/bench_mappo.py --steps 32 --updates 3 $ARGS 2>&1 | tail -1 | tee -a gpurun_out/mappo_$TAG.log
done
