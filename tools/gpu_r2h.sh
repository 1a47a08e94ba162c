#!/bin/bash
# Round-2 GPU call H: chain tests again, rules-bound variants at 32 envs/CTA in chains, bench with chained timing.
mkdir -p gpurun_out
TAG=${1:-r02h}
T0=$SECONDS; timeout 600 python -m pytest tests/test_gpu_step_chain.py tests/test_gpu_parity.py::test_cuda_replay_of_a_reference_game_and_recorder_round_trip -x -q > gpurun_out/pytest_chain_$TAG.log 2>&1; echo "pytest rc=$? wall $((SECONDS-T0)) s"
tail -15 gpurun_out/pytest_chain_$TAG.log
echo "== rules32"; PM_LIB_PATH=$PWD/build_variants/lib_rules32.so timeout 300 python tools/run_variants.py --chain 50 --reps 200 --small 0 --only views,bots,bots+views 2>&1 | grep -v "^{" | tee -a gpurun_out/variants_$TAG.log
T0=$SECONDS; timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2>gpurun_out/bench_$TAG.err; echo "bench rc=$? wall $((SECONDS-T0)) s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02h.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'roofline',d['roofline']['frac'],d['roofline']['kernel_ms'],'sep',d['roofline']['separate_launches'], d['separate_launches'])
print({k:(v['value'],v['roofline_frac'],v.get('separate_launches')) for k,v in d['also'].items() if k in ('smallCapture','bloxCapture')})
print('e2e',d['e2e']['value'], d['clocks'])
PY
tail -3 gpurun_out/bench_$TAG.err
