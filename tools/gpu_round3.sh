#!/bin/bash
set -x
mkdir -p gpurun_out
TAG=${1:-r01e}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_gpu_$TAG.log
for c in 0 2 4; do PM_BENCH_CHUNKS=$c python bench.py --steps 100 --warmup 5 --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunks',$c,'e2e',d['e2e']['value'],'plain',d['e2e']['unpipelined_pm_step_host_agent_steps_per_s'],'value',d['value'])"; done
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
