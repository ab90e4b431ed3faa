#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/stem_check.py 2>&1 | tee gpurun_out/stem_check.log | cut -c1-1200
if grep -q "stem_check fails: 0" gpurun_out/stem_check.log; then
  timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
  timeout 300 python bench.py --no-cpu-baseline --ops-out gpurun_out/bench_ops_resnet50.json > gpurun_out/bench_resnet50.json 2> gpurun_out/bench_resnet50.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_resnet50.json").read().strip().splitlines()[-1]); print("resnet50", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"])
o=json.load(open("gpurun_out/bench_ops_resnet50.json"))
for r in o["ops"][:3]: print(r)
PY
fi
