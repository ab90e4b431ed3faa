#!/bin/bash
# session 5 round 6: EfficientNet: parity + first bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "effi" > gpurun_out/pytest_gpu_effi.log 2>&1; tail -15 gpurun_out/pytest_gpu_effi.log | cut -c1-1500
timeout 400 python bench.py --model efficientnet_b0 --steps 30 --ops-out gpurun_out/ops_efficientnet_b0.json > gpurun_out/bench_efficientnet_b0.json 2> gpurun_out/bench_efficientnet_b0.err
tail -c 1800 gpurun_out/bench_efficientnet_b0.json; tail -5 gpurun_out/bench_efficientnet_b0.err
python - <<PY
import json
o=json.load(open("gpurun_out/ops_efficientnet_b0.json"))
print(o["ms_per_step"], sum(r["ms"] for r in o["ops"]), sum(r["t_bound_ms"] for r in o["ops"]))
for r in sorted(o["ops"], key=lambda r:-r["ms"])[:25]: print(f'{r["ms"]:.4f} {r["t_bound_ms"]:.4f} {r["frac"]:.2f} {r["op"]}')
PY
