#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_nets.py -q -k "seres or senet or seblock" 2>&1 | tail -12
for f in 1 0; do
PCV_SE_FOLD=$f timeout 600 python bench.py --model seresnext50_32x4d --steps 30 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/r02_se_$f.json 2> gpurun_out/r02_se_$f.err; tail -c 200 gpurun_out/r02_se_$f.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_se_$f.json').read().strip().splitlines()[-1])
print('fold=$f VALUE', d['value'], d['ms_per_step'], 'sustained', d['sustained']['value'], d['parity'], d['roofline_step']['frac'])
PY
done
