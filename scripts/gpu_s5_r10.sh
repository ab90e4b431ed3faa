#!/bin/bash
# session 5 round 10: two CTAs per SM for narrow short-K layers (single-CTA igemm kernel)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log | cut -c1-900
MODELS="mobilenetv2_w1 mobilenetv3_large_w1 efficientnet_b0" REPS=1 bash scripts/gpu_ab.sh 2>&1 | tail -8
python - <<PY
import json
for m in ("mobilenetv2_w1",):
  for w in ("prev","new"):
    o=json.load(open(f"gpurun_out/ab_ops_{m}_{w}.json"))
    sel=[r for r in o["ops"] if r["op"].startswith("conv_tc ")]
    print(m, w, round(sum(r["ms"] for r in sel),4), [(r["ms"]) for r in sel[:9]])
PY
