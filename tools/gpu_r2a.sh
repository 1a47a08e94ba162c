#!/bin/bash
# Round-2 GPU call A: parity suite (incl. full-size tests), bench both arms, the trainer's kernel variants timed
# and captured by ncu (launch list + full set).  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
TAG=${1:-r02a}
nproc; free -g | head -2; nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
T0=$SECONDS; python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$? wall $((SECONDS-T0)) s"
tail -30 gpurun_out/pytest_gpu_$TAG.log
T0=$SECONDS; python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref arm rc=$? wall $((SECONDS-T0)) s"
cat gpurun_out/bench_ref_$TAG.json
T0=$SECONDS; python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$? wall $((SECONDS-T0)) s"
cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
python tools/run_variants.py > gpurun_out/variants_$TAG.log 2>&1; cat gpurun_out/variants_$TAG.log
CMD="python tools/run_variants.py --reps 3 --warm 2"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_variants_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pm_env_kernel -c 60 -f -o gpurun_out/prof_variants_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out | tail -12
