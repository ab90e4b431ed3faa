#!/bin/bash
# Same-box A/B of two library builds: pytorchcv_b200/libpcv_b200_prev.so (PCV_B200_LIB) against the in-tree build.
mkdir -p gpurun_out
MODELS=${MODELS:-"resnet50 mobilenetv2_w1 seresnext50_32x4d deeplabv3_resnetd50b_voc"}
for rep in $(seq 1 ${REPS:-2}); do
for m in $MODELS; do
  for which in prev new; do
    if [ $which = prev ]; then export PCV_B200_LIB=$PWD/pytorchcv_b200/libpcv_b200_prev.so; else unset PCV_B200_LIB; fi
    timeout 300 python bench.py --model $m --no-cpu-baseline --no-configs --steps 30 --ops-out gpurun_out/ab_ops_${m}_$which.json > gpurun_out/ab_${m}_$which.json 2> gpurun_out/ab_${m}_$which.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_${m}_$which.json").read().strip().splitlines()[-1]); print("$m $which", d["value"], d["ms_per_step"], d["roofline_step"]["frac"])
except Exception as e: print("$m $which failed", e); print(open("gpurun_out/ab_${m}_$which.err").read()[-800:])
PY
  done
done
done
