#!/bin/bash
# fp32 tier on the tensor cores: kernel cases, fp32 net parity tests, resnet18 bs8 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "f32x3 or simt_f32 or dw3x3_s1_f32" 2>&1 | tail -30 > gpurun_out/r02_f32_kernels.log; tail -4 gpurun_out/r02_f32_kernels.log
timeout 1500 python -m pytest tests/test_gpu_nets.py -q -k "fp32 or f32" 2>&1 | tail -40 > gpurun_out/r02_f32_nets.log; tail -6 gpurun_out/r02_f32_nets.log
timeout 600 python bench.py --model resnet18 --steps 50 --warmup 5 --no-configs > gpurun_out/r02_f32_bench_resnet18.json 2> gpurun_out/r02_f32_bench_resnet18.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r02_f32_bench_resnet18.err
cp gpurun_out/bench_ops.json gpurun_out/r02_f32_bench_ops_resnet18.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_f32_bench_resnet18.json').read().strip().splitlines()[-1])
print('VALUE', d['value'], d['ms_per_step'], 'sustained', d['sustained']['value'], 'e2e', d['e2e']['value'], d['parity'])
for o in json.load(open('gpurun_out/bench_ops.json'))['ops']: print(f"{o['op']:75s} {o['ms']:.4f}")
PY
