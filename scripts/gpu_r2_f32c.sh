#!/bin/bash
mkdir -p gpurun_out
for ch in 4 8; do
echo "=== chunk $ch"
PCV_F3_CHUNK=$ch timeout 600 python -m pytest tests/test_gpu_nets.py -q -k "fp32 and seresnext" 2>&1 | grep -E "AssertionError: \(|assert 0\.|passed|failed" | head -8
PCV_F3_CHUNK=$ch timeout 300 python bench.py --model resnet18 --steps 50 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/r02_f32c_$ch.json 2> gpurun_out/r02_f32c_$ch.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_f32c_$ch.json').read().strip().splitlines()[-1])
print('chunk $ch VALUE', d['value'], d['ms_per_step'], d['parity']['rel_err'])
PY
done
