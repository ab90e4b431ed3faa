#!/bin/bash
# session 5 final: smoke, the contract bench line (default flags), the other four configs, the reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader; lscpu | grep -E "Model name|^CPU\(s\)" 
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/final_resnet50.json 2> gpurun_out/final_resnet50.err; tail -c 2500 gpurun_out/final_resnet50.json; tail -3 gpurun_out/final_resnet50.err
cp gpurun_out/bench_ops.json gpurun_out/final_ops_resnet50.json
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log | cut -c1-600
for m in resnet18 mobilenetv2_w1 seresnext50_32x4d deeplabv3_resnetd50b_voc efficientnet_b0 mobilenetv3_large_w1; do
  extra=""; [ $m = resnet18 ] && extra="--dtype fp32 --batch 8"
  timeout 600 python bench.py --model $m $extra --steps 50 --ops-out gpurun_out/final_ops_$m.json > gpurun_out/final_$m.json 2> gpurun_out/final_$m.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/final_$m.json").read().strip().splitlines()[-1])
    print("$m", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], d["roofline"]["kernel"][:50], "step", d["roofline_step"]["frac"], "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], d["clocks"])
except Exception as e: print("$m failed", e); print(open("gpurun_out/final_$m.err").read()[-600:])
PY
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err; cat gpurun_out/final_reference.json | cut -c1-900
