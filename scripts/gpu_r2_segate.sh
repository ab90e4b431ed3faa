#!/bin/bash
# SE scale fused into conv3's epilogue: parity tests of the SE families, then SE-ResNeXt-50 bench with and without (same box)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_kernels.py -q -x -k "se or SE or senet or benchmarked" 2>&1 | tail -4
for rep in 1 2; do
for f in 0 1; do
  PCV_SE_GATE_FUSE=$f timeout 600 python bench.py --model seresnext50_32x4d --no-cpu-baseline --no-configs --steps 30 --warmup 5 --ops-out gpurun_out/segate_ops_$f.json > gpurun_out/segate_bench_$f.json 2> gpurun_out/segate_bench_$f.err
  python - <<PY
import json
try:
    r=json.loads(open("gpurun_out/segate_bench_$f.json").read().strip().splitlines()[-1])
    print("se_gate_fuse=$f", r["value"], r["ms_per_step"], "sustained", r["sustained"]["value"], "e2e", r.get("e2e",{}).get("value"), r["parity"]["rel_err"], r["parity"]["top1_equal"], r["roofline_step"]["frac"], r["clocks"]["sm_mhz"])
except Exception as e:
    print("se_gate_fuse=$f failed", e); print(open("gpurun_out/segate_bench_$f.err").read()[-2000:])
PY
done
done
python - <<'PY'
import json
d=json.load(open("gpurun_out/segate_ops_1.json"))
for o in d["ops"][:22]: print("  ", o["op"][:100], round(o["ms"]*1000,1), o.get("frac"))
PY
