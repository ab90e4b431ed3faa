#!/bin/bash
# Session-4 state check: GPU parity tests, bench on all configs (per-op timings kept).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --ops-out gpurun_out/bench_ops_resnet50.json > gpurun_out/bench_resnet50.json 2> gpurun_out/bench_resnet50.err; tail -c 1500 gpurun_out/bench_resnet50.json
for m in mobilenetv2_w1 seresnext50_32x4d deeplabv3_resnetd50b_voc resnet18; do
  timeout 300 python bench.py --model $m --no-cpu-baseline --steps 30 --ops-out gpurun_out/bench_ops_$m.json > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$m.json").read().strip().splitlines()[-1]); print("$m", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline_step"])
except Exception as e: print("$m failed", e); print(open("gpurun_out/bench_$m.err").read()[-1500:])
PY
done
echo done
