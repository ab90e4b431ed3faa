#!/bin/bash
# two GPUs: the device-initiated peer all-gather test + a short weak / strong scaling bench (peer vs NCCL exchange)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_n2_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -15 > gpurun_out/r02_n2_pytest.log; tail -3 gpurun_out/r02_n2_pytest.log
for ex in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --exchange $ex > gpurun_out/r02_n2_bench_$ex.json 2> gpurun_out/r02_n2_bench_$ex.err
echo "bench $ex rc=$?"; tail -c 600 gpurun_out/r02_n2_bench_$ex.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 --scaling strong > gpurun_out/r02_n2_bench_strong.json 2> gpurun_out/r02_n2_bench_strong.err
python - <<'PY'
import json
for n in ("peer","nccl","strong"):
    try:
        d=json.loads(open(f'gpurun_out/r02_n2_bench_{n}.json').read().strip().splitlines()[-1])
        print(n, 'value', d['value'], 'ms', d['ms_per_step'], 'sustained', d['sustained']['value'], 'e2e', d['e2e']['value'], 'e2e_f32', d['e2e_f32']['value'], d['config']['parallelism'], d['parity']['rel_err'])
    except Exception as e: print(n, 'ERR', e)
PY
