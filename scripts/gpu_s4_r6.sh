#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/gpu_diag.py --timeout 60 --only seg_ --out gpurun_out/diag_seg.jsonl 2>&1 | cut -c1-400
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log | cut -c1-600
for m in resnet50 mobilenetv2_w1 seresnext50_32x4d deeplabv3_resnetd50b_voc; do
  timeout 300 python bench.py --model $m --no-cpu-baseline --steps 30 --ops-out gpurun_out/bench_ops_$m.json > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$m.json").read().strip().splitlines()[-1]); print("$m", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline_step"]["frac"])
except Exception as e: print("$m failed", e); print(open("gpurun_out/bench_$m.err").read()[-1500:])
PY
done
