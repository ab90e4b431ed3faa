#!/bin/bash
# gated dual-source conv with both TMEM loads issued up front: parity tests, then SE-ResNeXt-50 new build vs previous build
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_nets.py -q -x -m gpu -k "se_gate_in_conv3" 2>&1 | tail -3
m=seresnext50_32x4d
for which in new prev; do
  if [ $which = prev ]; then export PCV_B200_LIB=$PWD/pytorchcv_b200/libpcv_b200_prev.so; else unset PCV_B200_LIB; fi
  timeout 100 python bench.py --model $m --no-cpu-baseline --no-configs --steps 30 --ops-out gpurun_out/dualse2_ops_$which.json > gpurun_out/dualse2_$which.json 2> gpurun_out/dualse2_$which.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/dualse2_$which.json").read().strip().splitlines()[-1]); print("$which", d["value"], d["ms_per_step"], d["parity"]["rel_err"], d["clocks"]["sm_mhz"])
    for o in json.load(open("gpurun_out/dualse2_ops_$which.json"))["ops"]:
        if "*gate" in o["op"] and "+1x1" in o["op"]: print("   ", o["ms"], o["op"])
except Exception as e: print("$which failed", e); print(open("gpurun_out/dualse2_$which.err").read()[-800:])
PY
done
