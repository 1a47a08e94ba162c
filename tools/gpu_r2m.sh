#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02z}
export PM_LIB_PATH=$PWD/build_variants/lib_stag.so
for L in tinyCapture smallCapture bloxCapture; do
  timeout 300 python tools/run_variants.py --layout $L --reps 200 --small 0 --only plain 2>&1 | grep -v "^{\|^lib" | sed "s/^/stag /" | tee -a gpurun_out/variants_$TAG.log
done
timeout 300 python tools/run_variants.py --envs 131072 --small 49152 --reps 100 --only plain 2>&1 | grep -v "^{\|^lib" | sed "s/^/stag /" | tee -a gpurun_out/variants_$TAG.log
timeout 600 python -m pytest tests/test_gpu_full_size.py -m gpu -x -q -k "config2 or config3" 2>&1 | tail -3
unset PM_LIB_PATH
timeout 300 python tools/run_variants.py --envs 131072 --small 49152 --reps 100 --only plain 2>&1 | grep -v "^{\|^lib" | sed "s/^/default /" | tee -a gpurun_out/variants_$TAG.log
