#!/bin/bash
# session 5 round 1: new stem+pool epilogue / pool pass, F2FP.RELU epilogues, shared-space staging accesses
mkdir -p gpurun_out
timeout 600 python scripts/stem_check.py 2>&1 | tee gpurun_out/stem_check.log | cut -c1-300 | grep -v '"ok": true'
grep -c '"ok": true' gpurun_out/stem_check.log
if grep -q "stem_check fails: 0" gpurun_out/stem_check.log; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log | cut -c1-600
fi
MODELS="resnet50 mobilenetv2_w1" bash scripts/gpu_ab.sh
for d in 1 2 16 18 8; do
  PCV_STEM_DBG=$d timeout 200 python bench.py --model resnet50 --no-cpu-baseline --steps 10 --graph 0 --ops-out gpurun_out/stem_dbg$d.json > /dev/null 2> gpurun_out/stem_dbg$d.err
  python - <<PY
import json
o=json.load(open("gpurun_out/stem_dbg$d.json")); print("dbg $d", o["ops"][0]["ms"], o["ops"][0]["op"][:40])
PY
done
python - <<PY
import json
for w in ("prev","new"):
    o=json.load(open(f"gpurun_out/ab_ops_resnet50_{w}.json"))
    print(w, [ (r["ms"]) for r in o["ops"][:6]], sum(r["ms"] for r in o["ops"]))
PY
