#!/bin/bash
# session 5 round 8: MobileNetV3: parity + first bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "mobilenetv3 or mnv3" > gpurun_out/pytest_gpu_mnv3.log 2>&1; tail -12 gpurun_out/pytest_gpu_mnv3.log | cut -c1-1500
for m in mobilenetv3_large_w1 efficientnet_b0; do
timeout 400 python bench.py --model $m --steps 30 --no-cpu-baseline --ops-out gpurun_out/ops_$m.json > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$m.json").read().strip().splitlines()[-1]); print("$m", d["value"], d["ms_per_step"], d["roofline_step"]["frac"])
o=json.load(open("gpurun_out/ops_$m.json"))
for r in sorted(o["ops"], key=lambda r:-r["ms"])[:12]: print(f'   {r["ms"]:.4f} {r["t_bound_ms"]:.4f} {r["frac"]:.2f} {r["op"]}')
PY
done
