#!/bin/bash
# Round-2 evidence run on a B200 box (via gpurun): the gpu test suite, smoke, both bench arms, every step-kernel
# variant as separate launches and in pm_step_chain runs, the ncu launch list of the bench command and one full ncu
# capture per variant (raw page + per-segment and per-line summaries).  Outputs land in gpurun_out/; the files quoted
# in DESIGN.md are copied to profiles/ by hand.
mkdir -p gpurun_out
TAG=${1:-r02}
T0=$SECONDS; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$? wall $((SECONDS-T0)) s"
tail -4 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
T0=$SECONDS; timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2>gpurun_out/bench_ref_$TAG.err; echo "ref rc=$? wall $((SECONDS-T0)) s"
T0=$SECONDS; timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2>gpurun_out/bench_$TAG.err; echo "bench rc=$? wall $((SECONDS-T0)) s"
tail -c 3000 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
for L in tinyCapture smallCapture bloxCapture; do
  timeout 300 python tools/run_variants.py --layout $L --chain 50 --reps 200 --small $([ $L = tinyCapture ] && echo 1024 || echo 0) 2>&1 | grep -v "^{" | tee -a gpurun_out/variants_$TAG.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 40 --warmup 3 --no-extras > gpurun_out/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
bash tools/ncu_variants.sh $TAG > gpurun_out/ncu_variants_$TAG.log 2>&1
for V in plain views views+obs bots bots+views; do [ -f gpurun_out/prof_${V}_$TAG.ncu-rep ] && python tools/ncu_segments.py gpurun_out/prof_${V}_$TAG.ncu-rep; done 2>&1 | tee gpurun_out/ncu_segments_$TAG.txt
