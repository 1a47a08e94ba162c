#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r03e}
T0=$SECONDS; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$? wall $((SECONDS-T0)) s"
tail -3 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
T0=$SECONDS; timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2>gpurun_out/bench_ref_$TAG.err; echo "ref rc=$? wall $((SECONDS-T0)) s"
T0=$SECONDS; timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2>gpurun_out/bench_$TAG.err; echo "bench rc=$? wall $((SECONDS-T0)) s"
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print('value',d['value'],'roofline',d['roofline']['frac'],d['roofline']['write_only_reference']['frac_of_memset'],'sep',d['roofline']['separate_launches']['frac'])
print('e2e',d['e2e']['value'],'chain',d['e2e_chain']['value'])
print({k:(v['value'],v['roofline_frac']) for k,v in d['also'].items() if k in ('smallCapture','bloxCapture')})
print(json.dumps(d['also']['mappo'])[:1200])
print(json.dumps(d['also']['bloxCapture_envs_sweep']))
PY
