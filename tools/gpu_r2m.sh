#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02w}
timeout 600 python -m pytest tests/test_gpu_step_chain.py -m gpu -x -q 2>&1 | tail -3
for L in tinyCapture smallCapture bloxCapture; do
  timeout 300 python tools/run_variants.py --layout $L --chain 50 --reps 200 --small 16384 --only plain 2>&1 | grep -v "^{\|^lib" | tee -a gpurun_out/variants_$TAG.log
done
T0=$SECONDS; timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2>gpurun_out/bench_$TAG.err; echo "bench rc=$? wall $((SECONDS-T0)) s"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print('value',d['value'],'roofline',d['roofline']['frac'],d['roofline']['write_only_reference']['frac_of_memset'],'sep',d['roofline']['separate_launches']['frac'])
print('e2e',d['e2e']['value'],'chain',d['e2e_chain']['value'])
print({k:(v['value'],v['roofline_frac']) for k,v in d['also'].items() if k in ('smallCapture','bloxCapture')})
PY
