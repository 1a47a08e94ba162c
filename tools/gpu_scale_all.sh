#!/bin/bash
# weak-scaling sweep of bench.py (and the trainer at the largest N) on one box
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
for N in 1 2 4 8; do
  [ $N -gt $(nvidia-smi -L | wc -l) ] && break
  if [ $N -eq 1 ]; then python bench.py --steps 200 --warmup 10 --no-extras > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err; fi
  echo "N=$N rc=$?"; python -c "import json,sys; d=json.loads([l for l in open('gpurun_out/scale_${TAG}_n$N.json') if l.startswith('{')][-1]); print('N',d['n_gpus'],'value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])"
done
NG=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29540 tools/bench_mappo.py --envs 1024 --steps 32 --updates 2 --amp 1 > gpurun_out/mappo_${TAG}_n$NG.json 2> gpurun_out/mappo_${TAG}_n$NG.err; echo "mappo rc=$?"
cat gpurun_out/mappo_${TAG}_n$NG.json
