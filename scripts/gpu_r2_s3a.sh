#!/bin/bash
# session 3: dwpw tail blocks (prev vs new library on the same box) and the long-K wave term knob
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nets.py -q -x -k "fused_dw_pw" 2>&1 | tail -8
MODELS="mobilenetv2_w1" REPS=2 bash scripts/gpu_ab.sh
for kb in 0 16; do
  for rep in 1 2; do
  if [ $kb = 0 ]; then unset PCV_IGEMM2_WAVE_KB; else export PCV_IGEMM2_WAVE_KB=$kb; fi
  timeout 300 python bench.py --model resnet50 --no-cpu-baseline --no-configs --steps 30 --ops-out gpurun_out/wave_ops_$kb.json > gpurun_out/wave_$kb.json 2> gpurun_out/wave_$kb.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/wave_$kb.json").read().strip().splitlines()[-1]); print("resnet50 wave_kb=$kb", d["value"], d["ms_per_step"], d["sustained"]["value"], d["clocks"]["sm_mhz"])
except Exception as e: print("failed", e); print(open("gpurun_out/wave_$kb.err").read()[-800:])
PY
  done
done
python - <<'PY'
import json
for kb in (0,16):
    d=json.load(open(f"gpurun_out/wave_ops_{kb}.json"))
    for o in d["ops"]:
        if "2048->512" in o["op"] or "1024->512" in o["op"]: print(kb, o["op"], round(o["ms"]*1000,1))
for w in ("prev","new"):
    d=json.load(open(f"gpurun_out/ab_ops_mobilenetv2_w1_{w}.json"))
    print(w, d["ms_per_step"])
    for o in d["ops"][:12]: print("  ", o["op"][:90], round(o["ms"]*1000,1))
PY
