#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02n}
T0=$SECONDS; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$? wall $((SECONDS-T0)) s"
tail -6 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python tools/run_variants.py --layout bloxCapture --chain 50 --reps 200 --small 0 --only views,bots+views 2>&1 | grep -v "^{" | tee -a gpurun_out/variants_$TAG.log
