#!/bin/bash
# Round-2 GPU call D: control experiment for the geometry-variant libraries (clocks + host-drain indicator), encoder
# variants (256-bit stores, TMA bulk stores), static-plane elision, full bench.
set -x
mkdir -p gpurun_out
TAG=${1:-r02d}
cp pacman_marl_2025_b200/libpacman_b200.so build_variants/lib_control.so
echo "== default"; python tools/run_variants.py --small 0 2>&1 | grep -v "^{" | tee gpurun_out/variants_$TAG.log
for V in control st256 tma plainE64 botsE64; do
  echo "== $V"; PM_LIB_PATH=$PWD/build_variants/lib_$V.so python tools/run_variants.py --small 0 --only plain,bots,views 2>&1 | grep -v "^{" | sed "s/^/$V /" | tee -a gpurun_out/variants_$TAG.log
done
python tools/run_variants.py --small 0 --layout smallCapture --only plain,plain-elide,views 2>&1 | grep -v "^{" | tee -a gpurun_out/variants_$TAG.log
python tools/run_variants.py --small 0 --layout bloxCapture --only plain,plain-elide,views 2>&1 | grep -v "^{" | tee -a gpurun_out/variants_$TAG.log
T0=$SECONDS; python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$? wall $((SECONDS-T0)) s"
cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
T0=$SECONDS; python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$? wall $((SECONDS-T0)) s"
tail -6 gpurun_out/pytest_gpu_$TAG.log
