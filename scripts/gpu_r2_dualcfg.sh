#!/bin/bash
# ring / staging configuration sweep for the dual-source 1x1 conv (PCV_IGEMM2_DUAL_CFG = stages,ksub,nstg), ResNet-50 per-op table
mkdir -p gpurun_out
for cfg in default 5,1,4 4,1,6 6,1,2 3,2,2; do
  if [ $cfg = default ]; then unset PCV_IGEMM2_DUAL_CFG; else export PCV_IGEMM2_DUAL_CFG=$cfg; fi
  timeout 300 python bench.py --model resnet50 --no-cpu-baseline --no-configs --steps 30 --ops-out gpurun_out/dualcfg_ops_$cfg.json > gpurun_out/dualcfg_$cfg.json 2> gpurun_out/dualcfg_$cfg.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/dualcfg_$cfg.json").read().strip().splitlines()[-1]); print("cfg $cfg", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])
    for o in json.load(open("gpurun_out/dualcfg_ops_$cfg.json"))["ops"]:
        if "+1x1" in o["op"]: print("   ", o["ms"], o["op"])
except Exception as e: print("cfg $cfg failed", e); print(open("gpurun_out/dualcfg_$cfg.err").read()[-1500:])
PY
done
