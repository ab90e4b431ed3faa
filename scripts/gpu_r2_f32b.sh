#!/bin/bash
mkdir -p gpurun_out
for ch in 16 64 100000; do
PCV_F3_CHUNK=$ch timeout 300 python bench.py --model resnet18 --steps 50 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/r02_f32b_$ch.json 2> gpurun_out/r02_f32b_$ch.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_f32b_$ch.json').read().strip().splitlines()[-1])
print('chunk $ch VALUE', d['value'], d['ms_per_step'], d['parity']['rel_err'])
ops=json.load(open('gpurun_out/bench_ops.json'))['ops']
print(' '.join(f"{o['ms']:.3f}" for o in ops))
PY
done
grep -n "rel\|AssertionError: (" gpurun_out/r02_f32_nets.log | head
