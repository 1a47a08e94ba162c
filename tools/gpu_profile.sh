#!/bin/bash
# Run on the B200 box via gpurun: parity tests, bench, write-peak microbench, then the ncu launch list and one
# full capture of the step kernel.  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
T0=$SECONDS; python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$? wall $((SECONDS-T0)) s"
cat gpurun_out/bench_$TAG.json
[ -x tools/write_peak ] && ./tools/write_peak > gpurun_out/write_peak_$TAG.log 2>&1; cat gpurun_out/write_peak_$TAG.log
for l in tinyCapture bloxCapture; do python tools/time_rules_only.py $l; done 2>&1 | tee gpurun_out/rules_only_$TAG.log
CMD="python bench.py --steps 6 --warmup 3 --no-extras"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:pm_env_kernelILi0ELi0ELi0E -s 6 -c 1 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
