#!/bin/bash
# Round-2 GPU call F: the gpu test suite, smoke, both bench arms at HEAD.
mkdir -p gpurun_out
TAG=${1:-r02f}
T0=$SECONDS; python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$? wall $((SECONDS-T0)) s"
tail -6 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
T0=$SECONDS; python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2>gpurun_out/bench_ref_$TAG.err; echo "ref rc=$? wall $((SECONDS-T0)) s"
T0=$SECONDS; python bench.py > gpurun_out/bench_$TAG.json 2>gpurun_out/bench_$TAG.err; echo "bench rc=$? wall $((SECONDS-T0)) s"
tail -c 7000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
python tools/run_variants.py 2>&1 | grep -v "^{" | tee gpurun_out/variants_$TAG.log
