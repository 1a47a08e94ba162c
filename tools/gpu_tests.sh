#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-t}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/pytest_gpu_$TAG.log
