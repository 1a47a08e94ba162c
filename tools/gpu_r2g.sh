#!/bin/bash
# Round-2 GPU call G: pm_step_chain - parity tests, then timing of chained vs separate launches.
mkdir -p gpurun_out
TAG=${1:-r02g}
T0=$SECONDS; timeout 600 python -m pytest tests/test_gpu_step_chain.py tests/test_gpu_parity.py::test_cuda_replay_of_a_reference_game_and_recorder_round_trip -x -q > gpurun_out/pytest_chain_$TAG.log 2>&1; echo "pytest rc=$? wall $((SECONDS-T0)) s"
tail -15 gpurun_out/pytest_chain_$TAG.log
for L in tinyCapture smallCapture bloxCapture; do
  timeout 300 python tools/run_variants.py --layout $L --chain 50 --reps 200 --small 0 2>&1 | grep -v "^{" | tee -a gpurun_out/variants_$TAG.log
done
timeout 120 python tools/run_variants.py --chain 32 --reps 128 --envs 1024 --small 4096 2>&1 | grep -v "^{" | tee -a gpurun_out/variants_$TAG.log
