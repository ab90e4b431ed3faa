#!/bin/bash
# pcv_conv1x1_dual: parity tests, then a same-box A/B of PCV_DUAL_IDENTITY=0/1 on ResNet-50 and DeepLabv3.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_nets.py -q -x -m gpu -k "projection_shortcut or resnet50 or deeplab or benchmarked" 2>&1 | tail -15
for rep in 1 2; do
for m in ${MODELS:-resnet50 deeplabv3_resnetd50b_voc}; do
  for dual in 0 1; do
    PCV_DUAL_IDENTITY=$dual timeout 300 python bench.py --model $m --no-cpu-baseline --no-configs --steps 30 --ops-out gpurun_out/dual_ops_${m}_$dual.json > gpurun_out/dual_${m}_$dual.json 2> gpurun_out/dual_${m}_$dual.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/dual_${m}_$dual.json").read().strip().splitlines()[-1]); print("$m dual=$dual", d["value"], d["ms_per_step"], d["roofline_step"]["frac"], d.get("parity"), d["clocks"]["sm_mhz"])
except Exception as e: print("$m dual=$dual failed", e); print(open("gpurun_out/dual_${m}_$dual.err").read()[-1500:])
PY
  done
done
done
