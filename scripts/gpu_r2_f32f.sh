#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 8 15; do
PCV_F3_DBG=$dbg PCV_F3_CHUNK=100000 timeout 300 python bench.py --model resnet18 --steps 30 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/r02_f32f_$dbg.json 2> gpurun_out/r02_f32f_$dbg.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_f32f_$dbg.json').read().strip().splitlines()[-1])
ops=json.load(open('gpurun_out/bench_ops.json'))['ops']
print('dbg=$dbg ms', d['ms_per_step'], ' '.join(f"{o['ms']:.3f}" for o in ops))
PY
done
