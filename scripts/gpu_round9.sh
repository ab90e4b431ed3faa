#!/bin/bash
mkdir -p gpurun_out
timeout 100 python scripts/profile_ops.py --set mobilenet --only dw3 > gpurun_out/profile_ops7.log 2>&1; cat gpurun_out/profile_ops7.log
timeout 100 python scripts/profile_ops.py --set mobilenet --only dw3 --reps 10 2>&1 | head -3
for g in 0 1; do
timeout 200 python bench.py --no-cpu-baseline --graph $g --steps 50 --ops-out gpurun_out/bench_ops_resnet50_g$g.json > gpurun_out/bench_resnet50_g$g.json 2> gpurun_out/bench_resnet50.err; tail -c 300 gpurun_out/bench_resnet50.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_resnet50_g$g.json").read().strip().splitlines()[-1]); print("resnet50 graph=$g", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline_step"]["frac"], d["clocks"])
PY
done
timeout 200 python bench.py --model mobilenetv2_w1 --no-cpu-baseline --steps 50 --ops-out gpurun_out/bench_ops_mobilenetv2_w1.json > gpurun_out/bench_mobilenetv2_w1.json 2>gpurun_out/bench_mobilenetv2_w1.err;  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_mobilenetv2_w1.json").read().strip().splitlines()[-1]); print("mobilenet", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline_step"])
PY
