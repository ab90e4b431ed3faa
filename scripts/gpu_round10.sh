#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
for pdl in 1 0; do
PCV_PDL=$pdl timeout 200 python bench.py --no-cpu-baseline --steps 50 --ops-out gpurun_out/bench_ops_resnet50_pdl$pdl.json > gpurun_out/bench_resnet50_pdl$pdl.json 2> gpurun_out/bench_resnet50.err; tail -c 300 gpurun_out/bench_resnet50.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_resnet50_pdl$pdl.json").read().strip().splitlines()[-1]); print("resnet50 pdl=$pdl", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline_step"]["frac"], d["clocks"])
PY
done
PCV_PDL=1 timeout 200 python bench.py --no-cpu-baseline --steps 50 --graph 0 > gpurun_out/bench_resnet50_pdl1_g0.json 2> gpurun_out/bench_resnet50.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_resnet50_pdl1_g0.json").read().strip().splitlines()[-1]); print("resnet50 pdl=1 graph=0", d["value"], d["ms_per_step"], d["e2e"]["value"])
PY
