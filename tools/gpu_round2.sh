#!/bin/bash
set -x
mkdir -p gpurun_out
TAG=${1:-r01d}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
python tools/bench_mappo.py --envs 1024 --steps 32 --updates 2 > gpurun_out/mappo_$TAG.json 2> gpurun_out/mappo_$TAG.err; echo "mappo rc=$?"
cat gpurun_out/mappo_$TAG.json; tail -5 gpurun_out/mappo_$TAG.err
python tools/bench_mappo.py --envs 1024 --steps 32 --updates 2 --amp 1 > gpurun_out/mappo_amp_$TAG.json 2>> gpurun_out/mappo_$TAG.err; echo "mappo amp rc=$?"
cat gpurun_out/mappo_amp_$TAG.json
