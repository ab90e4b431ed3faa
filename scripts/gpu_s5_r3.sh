#!/bin/bash
# session 5 round 3: single-CTA igemm epilogue, SE excite kernels, bilinear rows kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log | cut -c1-900
REPS=1 bash scripts/gpu_ab.sh 2>&1 | tail -8
python - <<PY
import json
for m,pat in (("mobilenetv2_w1","conv_tc "),("seresnext50_32x4d","se_excite"),("deeplabv3_resnetd50b_voc","bilinear"),("seresnext50_32x4d","gavgpool")):
  for w in ("prev","new"):
    o=json.load(open(f"gpurun_out/ab_ops_{m}_{w}.json"))
    sel=[r for r in o["ops"] if pat in r["op"]]
    print(m, pat, w, round(sum(r["ms"] for r in sel),4), "of", round(sum(r["ms"] for r in o["ops"]),4), [r["ms"] for r in sel[:8]])
PY
