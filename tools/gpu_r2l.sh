#!/bin/bash
# Round-2 GPU call L: chain host-mapped test, both bench arms, launch list, ncu of all step variants (final kernels)
TAG=${1:-r02l}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_step_chain.py -x -q 2>&1 | tail -4
T0=$SECONDS; timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2>gpurun_out/bench_ref_$TAG.err; echo "ref rc=$? wall $((SECONDS-T0)) s"
T0=$SECONDS; timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2>gpurun_out/bench_$TAG.err; echo "bench rc=$? wall $((SECONDS-T0)) s"
tail -c 9000 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 40 --warmup 3 --no-extras > gpurun_out/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
bash tools/ncu_variants.sh $TAG > gpurun_out/ncu_variants_$TAG.log 2>&1
for V in plain views views+obs bots bots+views; do [ -f gpurun_out/prof_${V}_$TAG.ncu-rep ] && python tools/ncu_segments.py gpurun_out/prof_${V}_$TAG.ncu-rep; done 2>&1 | tee gpurun_out/ncu_segments_$TAG.txt
