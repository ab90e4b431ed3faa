#!/bin/bash
mkdir -p gpurun_out
for m in mobilenetv2_w1 seresnext50_32x4d deeplabv3_resnetd50b_voc resnet18; do
  timeout 200 python bench.py --model $m --no-cpu-baseline --steps 30 --ops-out gpurun_out/bench_ops_$m.json > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$m.json").read().strip().splitlines()[-1]); print("$m", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline_step"])
except Exception as e: print("$m failed", e); print(open("gpurun_out/bench_$m.err").read()[-1500:])
PY
done
