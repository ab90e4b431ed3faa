#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/stem_check.py 2>&1 | tee gpurun_out/stem_check.log | cut -c1-700 | grep -v '"ok": true' 
grep -c '"ok": true' gpurun_out/stem_check.log
if grep -q "stem_check fails: 0" gpurun_out/stem_check.log; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log | cut -c1-600
  MODELS="resnet50 seresnext50_32x4d" bash scripts/gpu_ab.sh
  python - <<PY
import json
o=json.load(open("gpurun_out/ab_ops_resnet50_new.json"))
for r in o["ops"][:3]+o["ops"][-2:]: print(r)
PY
fi
