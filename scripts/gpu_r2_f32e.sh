#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "f32x3" 2>&1 | tail -4
timeout 1200 python -m pytest tests/test_gpu_nets.py -q -k "fp32 and (resnet18 or resnet50 or mobilenetv2 or benchmarked or seresnext)" 2>&1 | tail -6
for bn in 128 64; do
PCV_F3_BN=$bn timeout 300 python bench.py --model resnet18 --steps 50 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/r02_f32e_$bn.json 2> gpurun_out/r02_f32e_$bn.err; tail -c 200 gpurun_out/r02_f32e_$bn.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_f32e_$bn.json').read().strip().splitlines()[-1])
print('bn<=$bn VALUE', d['value'], d['ms_per_step'], 'sustained', d['sustained']['value'], d['parity']['rel_err'])
ops=json.load(open('gpurun_out/bench_ops.json'))['ops']
print(' '.join(f"{o['ms']:.3f}" for o in ops))
PY
done
