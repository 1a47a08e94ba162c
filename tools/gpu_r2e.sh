#!/bin/bash
# Round-2 GPU call E: variants with the corrected layout path, encoder variants, aligned elision, bots without the
# unrolled look-ahead, ncu of the final variants.
set -x
mkdir -p gpurun_out
TAG=${1:-r02e}
echo "== default"; python tools/run_variants.py 2>&1 | grep -v "^{" | tee gpurun_out/variants_$TAG.log
for L in smallCapture bloxCapture; do python tools/run_variants.py --small 0 --layout $L 2>&1 | grep -v "^{" | tee -a gpurun_out/variants_$TAG.log; done
for V in st256 tma plainE64 botsE64 botsc6; do
  echo "== $V"; PM_LIB_PATH=$PWD/build_variants/lib_$V.so python tools/run_variants.py --small 0 --only plain,bots,views,bots+views 2>&1 | grep -v "^{" | sed "s/^/$V /" | tee -a gpurun_out/variants_$TAG.log
done
T0=$SECONDS; python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$? wall $((SECONDS-T0)) s"
tail -6 gpurun_out/pytest_gpu_$TAG.log
for ARGS in "--fused 1 --attn auto" "--fused 1 --attn auto --amp 1"; do
  python tools/bench_mappo.py --steps 32 --updates 3 $ARGS 2>&1 | tail -1 | tee -a gpurun_out/mappo_$TAG.log
done
bash tools/ncu_variants.sh $TAG
