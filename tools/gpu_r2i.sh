#!/bin/bash
# Round-2 GPU call I: run-time mover index in the rules phase - full gpu suite + variant timings (separate and chained).
mkdir -p gpurun_out
TAG=${1:-r02i}
timeout 300 python tools/run_variants.py --chain 50 --reps 200 2>&1 | grep -v "^{" | tee gpurun_out/variants_$TAG.log
for L in smallCapture bloxCapture; do
  timeout 300 python tools/run_variants.py --layout $L --chain 50 --reps 200 --small 0 2>&1 | grep -v "^{" | tee -a gpurun_out/variants_$TAG.log
done
for l in tinyCapture bloxCapture; do timeout 120 python tools/time_rules_only.py $l 2>&1 | grep "us/launch" | tee -a gpurun_out/variants_$TAG.log; done
T0=$SECONDS; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$? wall $((SECONDS-T0)) s"
tail -6 gpurun_out/pytest_gpu_$TAG.log
