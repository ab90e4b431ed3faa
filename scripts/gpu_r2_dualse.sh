#!/bin/bash
# pcv_conv1x1_dual_se (gated dual-source conv, two TMEM accumulators): parity tests, then SE-ResNeXt-50 with the shortcut fusion
# on every stage (PCV_DUAL_GATE_MIN_HW=0) against none (a threshold no map reaches), per-op tables for the per-stage policy
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_nets.py -q -x -m gpu -k "se_gate_in_conv3 or (bf16_tier and se) or fp16_tier_random or seresnext_unit" 2>&1 | tail -12
m=seresnext50_32x4d
for hw in 100000000 0 0; do
  PCV_DUAL_GATE_MIN_HW=$hw timeout 200 python bench.py --model $m --no-cpu-baseline --no-configs --steps 30 --ops-out gpurun_out/dualse_ops_$hw.json > gpurun_out/dualse_$hw.json 2> gpurun_out/dualse_$hw.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/dualse_$hw.json").read().strip().splitlines()[-1]); print("min_hw=$hw", d["value"], d["ms_per_step"], d["roofline_step"]["frac"], d["parity"]["rel_err"], d["parity"]["top1_equal"], d["clocks"]["sm_mhz"])
    for o in json.load(open("gpurun_out/dualse_ops_$hw.json"))["ops"]:
        if "*gate" in o["op"] and ("+1x1" in o["op"]) or (" s2 " in o["op"] and "1x1" in o["op"][:14]) : print("   ", o["ms"], o["op"])
except Exception as e: print("min_hw=$hw failed", e); print(open("gpurun_out/dualse_$hw.err").read()[-1500:])
PY
done
