#!/bin/bash
# grouped 3x3 halo kernel with N = 16 diagonal sub-block MMAs: kernel + net parity, then SE-ResNeXt-50 with and without (same box)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nets.py -q -x -k "conv or resnext or senet or benchmarked" 2>&1 | tail -2
for rep in 1 2; do
for f in 0 1; do
  PCV_IGEMM3_SUB16=$f timeout 600 python bench.py --model seresnext50_32x4d --no-cpu-baseline --no-configs --steps 30 --warmup 5 --ops-out gpurun_out/sub16_ops_$f.json > gpurun_out/sub16_bench_$f.json 2> gpurun_out/sub16_bench_$f.err
  python - <<PY
import json
try:
    r=json.loads(open("gpurun_out/sub16_bench_$f.json").read().strip().splitlines()[-1])
    print("sub16=$f", r["value"], r["ms_per_step"], r["parity"]["rel_err"], r["parity"]["top1_equal"])
except Exception as e:
    print("sub16=$f failed", e); print(open("gpurun_out/sub16_bench_$f.err").read()[-1500:])
PY
done
done
python - <<'PY'
import json
for f in (0,1):
    d=json.load(open(f"gpurun_out/sub16_ops_{f}.json"))
    print(f, [ (o["op"][:52], round(o["ms"]*1000,1)) for o in d["ops"] if o["op"].startswith("conv_tc3")][::3])
PY
