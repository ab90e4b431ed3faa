#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 2 4 3 6 7; do
  PCV_IGEMM3_DBG=$dbg timeout 200 python bench.py --no-cpu-baseline --steps 10 --ops-out gpurun_out/ops3_dbg$dbg.json > /dev/null 2>&1
  python - <<PY
import json
o=json.load(open("gpurun_out/ops3_dbg$dbg.json")); r=[x for x in o["ops"] if x["op"].startswith("conv_tc3")]; print("dbg=$dbg", r[0]["ms"], r[0]["op"][:60], "|", r[3]["ms"], r[3]["op"][:60])
PY
done
