#!/bin/bash
# same-box comparison of an older tree (git worktree add -f _old <commit>; make -C _old/pytorchcv_b200/csrc) with the current one,
# secondary families included - a regression guard for changes to shared kernels
mkdir -p gpurun_out
for rep in 1 2; do
for m in efficientnet_b0 mobilenetv3_large_w1 resnet50 deeplabv3_resnetd50b_voc; do
  for which in old new; do
    if [ $which = old ]; then dir=_old; else dir=.; fi
    (cd $dir && timeout 300 python bench.py --model $m --no-cpu-baseline --no-configs --steps 30 --warmup 5 --ops-out /tmp/ops_$which.json > /tmp/on_$which.json 2> /tmp/on_$which.err)
    python - <<PY
import json
try:
    d=json.loads(open("/tmp/on_$which.json").read().strip().splitlines()[-1]); print("$m $which", d["value"], d["ms_per_step"], d["parity"]["rel_err"])
except Exception as e: print("$m $which failed", e); print(open("/tmp/on_$which.err").read()[-600:])
PY
  done
done
done
