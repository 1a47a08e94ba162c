#!/bin/bash
python tools/bench_mappo.py --envs 1024 --steps 32 --updates 3 --amp 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bf16 sps',d['value'],'rollout',d['rollout_s'],'update',d['update_s'])"
python tools/bench_mappo.py --envs 1024 --steps 32 --updates 3 --amp 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fp32 sps',d['value'],'rollout',d['rollout_s'],'update',d['update_s'])"
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k mappo 2>&1 | tail -3
