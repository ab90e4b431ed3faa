#!/bin/bash
# session 5 round 5: igemm3 MMA issue loop (one elect region per M-block: 36 MMAs back to back)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log | cut -c1-900
MODELS="resnet50 seresnext50_32x4d deeplabv3_resnetd50b_voc" REPS=1 bash scripts/gpu_ab.sh 2>&1 | tail -8
python - <<PY
import json
for m in ("resnet50","seresnext50_32x4d","deeplabv3_resnetd50b_voc"):
  for w in ("prev","new"):
    o=json.load(open(f"gpurun_out/ab_ops_{m}_{w}.json"))
    sel=[r for r in o["ops"] if "conv_tc3" in r["op"]]
    d={}
    for r in sel: d.setdefault(r["op"][8:60],[]).append(r["ms"])
    print(m, w, round(sum(r["ms"] for r in sel),4), {k:round(sum(v)/len(v),4) for k,v in d.items()})
PY
