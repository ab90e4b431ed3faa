#!/bin/bash
# session 5 round 2: role-split stem+pool, wide s2d ingest, narrow layers on the pair kernel
mkdir -p gpurun_out
timeout 600 python scripts/stem_check.py 2>&1 | tee gpurun_out/stem_check.log | cut -c1-300 | grep -v '"ok": true'
grep -c '"ok": true' gpurun_out/stem_check.log
if grep -q "stem_check fails: 0" gpurun_out/stem_check.log; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log | cut -c1-600
fi
MODELS="resnet50 mobilenetv2_w1" bash scripts/gpu_ab.sh 2>&1 | tail -4
PCV_IGEMM2_MIN_COUT=64 timeout 200 python bench.py --model mobilenetv2_w1 --no-cpu-baseline --steps 30 --ops-out gpurun_out/mnv2_min64_ops.json 2> /dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mnv2 MIN_COUT=64', d['value'], d['ms_per_step'])"
for d in 2 16 18 8; do
  PCV_STEM_DBG=$d timeout 200 python bench.py --model resnet50 --no-cpu-baseline --steps 10 --graph 0 --ops-out gpurun_out/stem_dbg$d.json > /dev/null 2> gpurun_out/stem_dbg$d.err
  python - <<PY
import json
o=json.load(open("gpurun_out/stem_dbg$d.json")); print("dbg $d", o["ops"][0]["ms"], o["ops"][0]["op"][:40])
PY
done
python - <<PY
import json
for m in ("resnet50","mobilenetv2_w1"):
  for w in ("prev","new"):
    o=json.load(open(f"gpurun_out/ab_ops_{m}_{w}.json"))
    print(m, w, [ (r["ms"]) for r in o["ops"][:6]], round(sum(r["ms"] for r in o["ops"]),4))
PY
