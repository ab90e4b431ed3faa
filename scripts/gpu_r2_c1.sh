#!/bin/bash
# round 2, call 1: full GPU test suite + the default bench line (all BASELINE configs) + a short reference-arm check
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_c1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py 2>&1 | tail -60 > gpurun_out/r02_c1_pytest.log
echo "pytest rc=$?" >> gpurun_out/r02_c1_pytest.log
tail -5 gpurun_out/r02_c1_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c1_bench.json 2> gpurun_out/r02_c1_bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r02_c1_bench.err
cp gpurun_out/bench_ops.json gpurun_out/r02_c1_bench_ops_resnet50.json 2>/dev/null
for m in resnet18 mobilenetv2_w1 seresnext50_32x4d deeplabv3_resnetd50b_voc; do cp gpurun_out/bench_ops_$m.json gpurun_out/r02_c1_bench_ops_$m.json 2>/dev/null; done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_c1_ref.json 2> gpurun_out/r02_c1_ref.err
echo "ref rc=$?"; cat gpurun_out/r02_c1_ref.json | head -c 600
python -c "
import json
d=json.loads(open('gpurun_out/r02_c1_bench.json').read().strip().splitlines()[-1])
print('VALUE', d['value'], 'sustained', d.get('sustained',{}).get('value'), 'e2e', d['e2e']['value'], 'e2e_f32', d['e2e_f32']['value'], 'parity', d['parity'])
print('roof_step', d['roofline_step'])
for c in d.get('configs', []): print(c.get('model'), c.get('value'), c.get('e2e'), c.get('roofline_step'), c.get('parity'), c.get('error'))
"
