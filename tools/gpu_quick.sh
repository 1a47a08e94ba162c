#!/bin/bash
mkdir -p gpurun_out
for l in tinyCapture smallCapture bloxCapture; do python tools/time_rules_only.py $l 2>&1 | grep "us/launch"; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -40
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_quick.json 2>gpurun_out/bench_quick.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print('value',d['value'],'frac',d['roofline']['frac'],'kernel_ms',d['roofline']['kernel_ms'],'e2e',d['e2e']['value'])
print({k:(v['value'],v['roofline_frac']) for k,v in d['also'].items() if k not in ('kernels','mappo')})
print({k:v.get('ms') for k,v in d['also']['kernels'].items()})
PY
