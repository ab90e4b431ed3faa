#!/bin/bash
# fused dw->pw: parity tests, then the MobileNetV2 bench with and without the fusion
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nets.py -q -x -k "fused_dw_pw" 2>&1 | tail -25 > gpurun_out/dwpw_tests.log
cat gpurun_out/dwpw_tests.log
for f in 1 0; do
  PCV_FUSE_DWPW=$f timeout 600 python bench.py --model mobilenetv2_w1 --steps 20 --warmup 5 --ops-out gpurun_out/dwpw_ops_$f.json > gpurun_out/dwpw_bench_$f.json 2> gpurun_out/dwpw_bench_$f.err
  python - <<PY
import json
try:
    r=json.loads(open("gpurun_out/dwpw_bench_$f.json").read().strip().splitlines()[-1])
    print("fuse=$f", r["value"], r["ms_per_step"], r.get("e2e",{}).get("value"), r.get("parity"), r.get("roofline",{}).get("frac"))
except Exception as e:
    print("fuse=$f failed", e); print(open("gpurun_out/dwpw_bench_$f.err").read()[-2000:])
PY
done
