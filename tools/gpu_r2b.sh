#!/bin/bash
# Round-2 GPU call B: tests that changed, variants with the new views encoder, attention formulations, MAPPO SPS with
# the fused pipeline, ncu of every variant.
set -x
mkdir -p gpurun_out
TAG=${1:-r02b}
T0=$SECONDS; python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$? wall $((SECONDS-T0)) s"
tail -25 gpurun_out/pytest_gpu_$TAG.log
python tools/run_variants.py > gpurun_out/variants_$TAG.log 2>&1; cat gpurun_out/variants_$TAG.log | grep -v "^{"
python tools/run_variants.py --layout smallCapture --small 0 2>&1 | grep -v "^{" | tee -a gpurun_out/variants_$TAG.log
python tools/run_variants.py --layout bloxCapture --small 0 2>&1 | grep -v "^{" | tee -a gpurun_out/variants_$TAG.log
python tools/bench_attention.py 2>&1 | tee gpurun_out/attention_$TAG.log
for ARGS in "--fused 0 --attn module" "--fused 1 --attn module" "--fused 0 --attn bmm" "--fused 1 --attn bmm" "--fused 1 --attn bmm --amp 1" "--fused 0 --attn module --amp 1"; do
  python tools/bench_mappo.py --steps 32 --updates 3 $ARGS 2>&1 | tail -1 | tee -a gpurun_out/mappo_$TAG.log
done
bash tools/ncu_variants.sh $TAG
