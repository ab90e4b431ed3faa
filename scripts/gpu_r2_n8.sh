#!/bin/bash
# N GPUs (argument): weak-scaling bench with the peer exchange, plus strong scaling
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_n${N}_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r02_n${N}_bench.json 2> gpurun_out/r02_n${N}_bench.err
echo "weak rc=$?"; tail -c 400 gpurun_out/r02_n${N}_bench.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 50 --warmup 5 --scaling strong > gpurun_out/r02_n${N}_bench_strong.json 2> gpurun_out/r02_n${N}_bench_strong.err
echo "strong rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 50 --warmup 5 --exchange nccl > gpurun_out/r02_n${N}_bench_nccl.json 2> gpurun_out/r02_n${N}_bench_nccl.err
python - <<PY
import json
for n in ("bench","bench_strong","bench_nccl"):
    try:
        d=json.loads(open(f'gpurun_out/r02_n${N}_{n}.json').read().strip().splitlines()[-1])
        print(n, 'value', d['value'], 'ms', d['ms_per_step'], 'sustained', d['sustained']['value'], 'e2e(u8)', d['e2e']['value'], 'e2e_f32', d['e2e_f32']['value'], d['config']['parallelism'], d['config']['numa'], d['parity']['rel_err'])
    except Exception as e: print(n, 'ERR', e)
PY
