#!/bin/bash
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
N=${2:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo "bench rc=$?"
cat gpurun_out/bench_${TAG}_n$N.json; tail -3 gpurun_out/bench_${TAG}_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_ref_${TAG}_n$N.json 2>&1; echo "ref rc=$?"
cat gpurun_out/bench_ref_${TAG}_n$N.json | tail -2
NCCL_DEBUG=WARN python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/bench_mappo.py --envs 1024 --steps 32 --updates 2 --amp 1 > gpurun_out/mappo_${TAG}_n$N.json 2> gpurun_out/mappo_${TAG}_n$N.err; echo "mappo rc=$?"
cat gpurun_out/mappo_${TAG}_n$N.json; tail -3 gpurun_out/mappo_${TAG}_n$N.err
