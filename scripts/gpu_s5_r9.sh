#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "dw5x5 or effi or mnv3 or mobilenetv3 or efficientnet" > gpurun_out/pytest_gpu_dw5.log 2>&1; tail -4 gpurun_out/pytest_gpu_dw5.log | cut -c1-800
for m in efficientnet_b0 mobilenetv3_large_w1; do
timeout 400 python bench.py --model $m --steps 30 --no-cpu-baseline --ops-out gpurun_out/ops_$m.json > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$m.json").read().strip().splitlines()[-1]); print("$m", d["value"], d["ms_per_step"], d["roofline_step"]["frac"])
o=json.load(open("gpurun_out/ops_$m.json"))
for r in sorted([r for r in o["ops"] if "5x5" in r["op"]], key=lambda r:-r["ms"])[:8]: print(f'   {r["ms"]:.4f} {r["t_bound_ms"]:.4f} {r["frac"]:.2f} {r["op"]}')
PY
done
