#!/bin/bash
# One gpurun call: GPU parity tests, bench (all configs), isolated op timings, ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --ops-out gpurun_out/bench_ops_resnet50.json > gpurun_out/bench_resnet50.json 2> gpurun_out/bench_resnet50.err; tail -c 1500 gpurun_out/bench_resnet50.json
timeout 300 python scripts/profile_ops.py --set all > gpurun_out/profile_ops.log 2>&1; cat gpurun_out/profile_ops.log
for m in mobilenetv2_w1 seresnext50_32x4d deeplabv3_resnetd50b_voc; do
  timeout 400 python bench.py --model $m --no-cpu-baseline --ops-out gpurun_out/bench_ops_$m.json > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$m.json").read().strip().splitlines()[-1]); print("$m", d["value"], d["ms_per_step"], d["roofline_step"])
except Exception as e: print("$m failed", e)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/launches_resnet50.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo done
