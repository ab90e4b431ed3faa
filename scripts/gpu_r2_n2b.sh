#!/bin/bash
# two GPUs, final build: the peer all-gather test and one short weak-scaling bench line (what the driver's scaling run launches)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_final_n2_bench.json 2> gpurun_out/r02_final_n2_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r02_final_n2_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_final_n2_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'sustained', d['sustained']['value'], 'e2e', d['e2e']['value'], d['config']['parallelism'], d['parity']['rel_err'], d['n_gpus'])
PY
