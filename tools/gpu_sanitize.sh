#!/bin/bash
# one sanitizer tool per call (see B200_PROFILING.md); small parity cases only
TOOL=${1:-memcheck}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fixture and (tiny or caps or odd15 or mini) or max_size or views_with_bots or env_ranges or bfs_distances or gae_bit" > gpurun_out/san_plain.log 2>&1 || { tail -20 gpurun_out/san_plain.log; exit 1; }
tail -2 gpurun_out/san_plain.log
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 9 --log-file gpurun_out/sanitizer_$TOOL.log python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fixture and (tiny or caps or odd15 or mini) or max_size or views_with_bots or env_ranges or bfs_distances or gae_bit" > gpurun_out/san_run_$TOOL.log 2>&1
echo "sanitizer rc=$?"
tail -3 gpurun_out/san_run_$TOOL.log
tail -8 gpurun_out/sanitizer_$TOOL.log
