#!/bin/bash
# session 5 round 7: fast swish-family epilogues, 5x5 depthwise through the TMA window kernel, EfficientNet-b0
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log | cut -c1-1500
MODELS="efficientnet_b0 mobilenetv2_w1 resnet50" REPS=1 bash scripts/gpu_ab.sh 2>&1 | tail -8
python - <<PY
import json
o=json.load(open("gpurun_out/ab_ops_efficientnet_b0_new.json"))
print(o["ms_per_step"], sum(r["ms"] for r in o["ops"]), sum(r["t_bound_ms"] for r in o["ops"]))
for r in sorted(o["ops"], key=lambda r:-r["ms"])[:22]: print(f'{r["ms"]:.4f} {r["t_bound_ms"]:.4f} {r["frac"]:.2f} {r["op"]}')
PY
