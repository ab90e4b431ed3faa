#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gpu_diag.py --timeout 40 --out gpurun_out/diag_all.jsonl 2>&1 | cut -c1-600 | grep -v '"ok": true'
timeout 120 python scripts/profile_ops.py --set resnet50 --reps 3 > gpurun_out/profile_ops8.log 2>&1; cat gpurun_out/profile_ops8.log
echo "== epilogue only"; PCV_IGEMM_DBG=3 PCV_IGEMM_HALO=0 timeout 100 python scripts/profile_ops.py --set resnet50 --reps 3 2>&1 | awk '{printf "%s %s | ", $1, $2} END {print ""}'
timeout 200 python bench.py --no-cpu-baseline --steps 50 --ops-out gpurun_out/bench_ops_resnet50.json > gpurun_out/bench_resnet50.json 2> gpurun_out/bench_resnet50.err; tail -c 300 gpurun_out/bench_resnet50.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_resnet50.json").read().strip().splitlines()[-1]); print("resnet50", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline_step"], d["clocks"])
PY
