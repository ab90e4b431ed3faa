#!/bin/bash
# 2-GPU check of the contract launch line (weak scaling, one all-gather of logits) + the reference arm under torchrun
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 1500 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
echo "reference arm under torchrun:"; cut -c1-300 gpurun_out/bench_ref_n2.json; tail -2 gpurun_out/bench_ref_n2.err
python - <<PY
import json
a=json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1]); b=json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print("N=1", a["value"], a["e2e"]["value"], a["clocks"]); print("N=2", b["value"], b["e2e"]["value"], b["clocks"], "eff", b["value"]/(2*a["value"]))
PY
