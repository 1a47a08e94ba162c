#!/bin/bash
# both bench arms under torchrun on every GPU of the box
mkdir -p gpurun_out
TAG=${1:-r02}
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
T0=$SECONDS; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --impl reference > gpurun_out/bench_ref_${TAG}_n$NG.json 2>gpurun_out/bench_ref_${TAG}_n$NG.err; echo "ref rc=$? wall $((SECONDS-T0)) s"
T0=$SECONDS; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $NG > gpurun_out/bench_${TAG}_n$NG.json 2>gpurun_out/bench_${TAG}_n$NG.err; echo "bench rc=$? wall $((SECONDS-T0)) s"
tail -5 gpurun_out/bench_${TAG}_n$NG.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_${TAG}_n$NG.json') if l.startswith('{')][-1])
print('n',d['n_gpus'],'value',d['value'],'roofline',d['roofline']['frac'],'sep',d['roofline']['separate_launches']['frac'], d['separate_launches']['value'])
print('e2e',d['e2e']['value'],'chain',d['e2e_chain']['value'], d['clocks'])
print(json.dumps(d['also'])[:1500])
PY
