#!/bin/bash
# End-of-round check on one B200: the gpu suite, smoke, both bench arms, and the variant table (separate / chained).
mkdir -p gpurun_out
TAG=${1:-final}
T0=$SECONDS; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$? wall $((SECONDS-T0)) s"
tail -3 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
T0=$SECONDS; timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2>gpurun_out/bench_ref_$TAG.err; echo "ref rc=$? wall $((SECONDS-T0)) s"
T0=$SECONDS; timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2>gpurun_out/bench_$TAG.err; echo "bench rc=$? wall $((SECONDS-T0)) s"
tail -3 gpurun_out/bench_$TAG.err
for L in tinyCapture smallCapture bloxCapture; do
  timeout 300 python tools/run_variants.py --layout $L --chain 50 --reps 200 --small $([ $L = tinyCapture ] && echo 1024 || echo 0) 2>&1 | grep -v "^{" | tee -a gpurun_out/variants_$TAG.log
done
