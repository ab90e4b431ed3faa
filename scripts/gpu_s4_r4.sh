#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 2 4 8 3 6 7; do
  PCV_STEM_DBG=$dbg timeout 200 python bench.py --no-cpu-baseline --steps 10 --ops-out gpurun_out/ops_dbg$dbg.json > /dev/null 2>&1
  python - <<PY
import json
o=json.load(open("gpurun_out/ops_dbg$dbg.json")); print("dbg=$dbg", o["ops"][0]["ms"], o["ops"][0]["op"][:40])
PY
done
