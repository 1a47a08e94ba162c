#!/bin/bash
# Round-2 GPU call C: changed tests, bots with the staged distance table, launch-geometry variants, trainer SPS.
set -x
mkdir -p gpurun_out
TAG=${1:-r02c}
T0=$SECONDS; python -m pytest tests/test_gpu_ppo_pipeline.py tests/test_gpu_parity.py -m gpu -x -q -k "ppo or adam or gather or trainer or pool or elision or baseline or bots or rng or views" > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$? wall $((SECONDS-T0)) s"
tail -15 gpurun_out/pytest_gpu_$TAG.log
echo "== default build"; python tools/run_variants.py 2>&1 | grep -v "^{" | tee gpurun_out/variants_$TAG.log
for L in smallCapture bloxCapture; do python tools/run_variants.py --layout $L --small 0 --only bots,bots+views 2>&1 | grep -v "^{" | tee -a gpurun_out/variants_$TAG.log; done
for V in botsE64 botsE128 botsE64c6 botsE32c6; do
  echo "== $V"; PM_LIB_PATH=$PWD/build_variants/lib_$V.so python tools/run_variants.py --small 0 --only bots,bots+views 2>&1 | grep -v "^{" | sed "s/^/$V /" | tee -a gpurun_out/variants_$TAG.log
done
for V in plainE64 plainE128; do
  echo "== $V"; PM_LIB_PATH=$PWD/build_variants/lib_$V.so python tools/run_variants.py --small 0 --only plain,views 2>&1 | grep -v "^{" | sed "s/^/$V /" | tee -a gpurun_out/variants_$TAG.log
  PM_LIB_PATH=$PWD/build_variants/lib_$V.so python tools/run_variants.py --small 0 --layout bloxCapture --only plain 2>&1 | grep -v "^{" | sed "s/^/$V /" | tee -a gpurun_out/variants_$TAG.log
done
for ARGS in "--fused 0 --attn module" "--fused 1 --attn bmm" "--fused 0 --attn bmm" "--fused 1 --attn module --amp 1"; do
  python tools/bench_mappo.py --steps 32 --updates 3 $ARGS 2>&1 | tail -1 | tee -a gpurun_out/mappo_$TAG.log
done
