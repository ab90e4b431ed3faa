#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/stem_check.py 2>&1 | tee gpurun_out/stem_check.log | cut -c1-300
for dbg in 0 2 1; do
  PCV_STEM_DBG=$dbg timeout 200 python bench.py --no-cpu-baseline --steps 10 --ops-out gpurun_out/ops_dbg$dbg.json > gpurun_out/bench_dbg$dbg.json 2>&1
  python - <<PY
import json
o=json.load(open("gpurun_out/ops_dbg$dbg.json")); print("dbg=$dbg", o["ops"][0]["ms"], o["ops"][0]["op"][:40], o["ms_per_step"])
PY
done
