#!/bin/bash
mkdir -p gpurun_out
T0=$SECONDS; python bench.py > gpurun_out/bench_${1:-x}.json 2> gpurun_out/bench_${1:-x}.err; echo "bench rc=$? wall $((SECONDS-T0)) s"
T0=$SECONDS; python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_${1:-x}.json 2>&1; echo "ref rc=$? wall $((SECONDS-T0)) s"; tail -c 600 gpurun_out/bench_ref_${1:-x}.json
