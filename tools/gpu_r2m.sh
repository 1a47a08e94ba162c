#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02x}
for V in b128c5 b128c6 b128c8; do
  for L in tinyCapture bloxCapture; do
    PM_LIB_PATH=$PWD/build_variants/lib_$V.so timeout 300 python tools/run_variants.py --layout $L --reps 100 --small 0 --only bots,bots+views 2>&1 | grep -v "^{\|^lib" | sed "s/^/$V /" | tee -a gpurun_out/variants_$TAG.log
  done
done
