#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02t}
for V in bots6 bots5; do
  for L in tinyCapture bloxCapture; do
    PM_LIB_PATH=$PWD/build_variants/lib_$V.so timeout 300 python tools/run_variants.py --layout $L --reps 100 --small 0 --only bots,bots+views 2>&1 | grep -v "^{\|^lib" | sed "s/^/$V /" | tee -a gpurun_out/variants_$TAG.log
  done
done
timeout 300 python tools/run_variants.py --chain 50 --reps 200 --only plain,bots,bots+views,plain-elide 2>&1 | grep -v "^{\|^lib" | sed "s/^/default /" | tee -a gpurun_out/variants_$TAG.log
timeout 300 python tools/run_variants.py --envs 262144 --small 16384 --chain 20 --reps 40 --only plain 2>&1 | grep -v "^{\|^lib" | sed "s/^/default /" | tee -a gpurun_out/variants_$TAG.log
