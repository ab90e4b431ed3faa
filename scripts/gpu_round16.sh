#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gpu_diag.py --timeout 40 --only group --out gpurun_out/diag_grp.jsonl 2>&1 | cut -c1-700
timeout 100 python scripts/profile_ops.py --set se --reps 3
timeout 200 python bench.py --model seresnext50_32x4d --no-cpu-baseline --steps 30 --ops-out gpurun_out/bench_ops_seresnext50_32x4d.json > gpurun_out/bench_seresnext50_32x4d.json 2> gpurun_out/bench_se.err; tail -c 400 gpurun_out/bench_se.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_seresnext50_32x4d.json").read().strip().splitlines()[-1]); print("seresnext", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline_step"])
PY
