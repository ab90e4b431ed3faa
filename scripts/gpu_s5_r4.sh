#!/bin/bash
# session 5 round 4: stem + max pool with the vertical max in registers (PW = 128: TMEM lane == column)
mkdir -p gpurun_out
timeout 900 python scripts/stem_check.py 2>&1 | tee gpurun_out/stem_check.log | cut -c1-400 | grep -v '"ok": true'
grep -c '"ok": true' gpurun_out/stem_check.log
if grep -q "stem_check fails: 0" gpurun_out/stem_check.log; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log | cut -c1-900
fi
MODELS="resnet50 seresnext50_32x4d" REPS=2 bash scripts/gpu_ab.sh 2>&1 | tail -8
for d in 2 16 18 8; do
  PCV_STEM_DBG=$d timeout 200 python bench.py --model resnet50 --no-cpu-baseline --steps 10 --graph 0 --ops-out gpurun_out/stem_dbg$d.json > /dev/null 2> gpurun_out/stem_dbg$d.err
  python - <<PY
import json
o=json.load(open("gpurun_out/stem_dbg$d.json")); print("dbg $d", o["ops"][0]["ms"], o["ops"][0]["op"][:60])
PY
done
python - <<PY
import json
for m in ("resnet50",):
  for w in ("prev","new"):
    o=json.load(open(f"gpurun_out/ab_ops_{m}_{w}.json"))
    print(m, w, [ (r["ms"], r["frac"], r["bound"]) for r in o["ops"][:2]], round(sum(r["ms"] for r in o["ops"]),4), o["ops"][0]["op"])
PY
