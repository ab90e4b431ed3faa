#!/bin/bash
# expansion -> dw -> pw fused kernel: parity tests, then the MobileNetV2 bench with and without it (same box), per-op tables
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_nets.py -q -x -k "test_fused_dw_pw" 2>&1 | tail -5
for rep in 1 2; do
for f in 0 1; do
  PCV_XDWPW_S1=${S1:-0} PCV_FUSE_XDWPW=$f timeout 600 python bench.py --model mobilenetv2_w1 --no-cpu-baseline --no-configs --steps 30 --warmup 5 --ops-out gpurun_out/xdwpw_ops_$f.json > gpurun_out/xdwpw_bench_$f.json 2> gpurun_out/xdwpw_bench_$f.err
  python - <<PY
import json
try:
    r=json.loads(open("gpurun_out/xdwpw_bench_$f.json").read().strip().splitlines()[-1])
    print("xdwpw=$f", r["value"], r["ms_per_step"], "sustained", r["sustained"]["value"], "e2e", r.get("e2e",{}).get("value"), r["parity"]["rel_err"], r["roofline_step"]["frac"])
except Exception as e:
    print("xdwpw=$f failed", e); print(open("gpurun_out/xdwpw_bench_$f.err").read()[-2000:])
PY
done
done
python - <<'PY'
import json
for f in (0,1):
    d=json.load(open(f"gpurun_out/xdwpw_ops_{f}.json"))
    print(f, d["ms_per_step"])
    for o in d["ops"][:14]: print("  ", o["op"][:100], round(o["ms"]*1000,1), o.get("frac"))
PY
