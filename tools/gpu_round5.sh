#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r01k}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for pair in 0 1; do python tools/bench_mappo.py --envs 1024 --steps 32 --updates 2 --amp 1 --pair $pair 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pair',$pair,'sps',d['value'],'rollout',d['rollout_s'],'update',d['update_s'])"; done
python tools/bench_mappo.py --envs 1024 --steps 32 --updates 2 --amp 1 --views 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('views 0 sps',d['value'],'rollout',d['rollout_s'],'update',d['update_s'])"
( time python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ) 2>&1 | grep real
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print('value',d['value'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'])
print(d['also']['mappo'])
PY
