#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "f32x3" 2>&1 | tail -3
for ks in 2 1 4; do
PCV_F3_KSUB=$ks timeout 300 python bench.py --model resnet18 --steps 50 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/r02_f32g_$ks.json 2> gpurun_out/r02_f32g_$ks.err; tail -c 200 gpurun_out/r02_f32g_$ks.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_f32g_$ks.json').read().strip().splitlines()[-1])
ops=json.load(open('gpurun_out/bench_ops.json'))['ops']
print('ksub=$ks VALUE', d['value'], d['ms_per_step'], d['parity']['rel_err'], ' '.join(f"{o['ms']:.3f}" for o in ops))
PY
done
timeout 1200 python -m pytest tests/test_gpu_nets.py -q -k "fp32" 2>&1 | tail -4
