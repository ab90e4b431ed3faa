#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "conv_block" 2>&1 | tail -3
timeout 600 python bench.py --steps 50 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/r02_ns_bench.json 2> gpurun_out/r02_ns_bench.err; tail -c 300 gpurun_out/r02_ns_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_ns_bench.json').read().strip().splitlines()[-1])
print('VALUE', d['value'], d['ms_per_step'], 'sustained', d['sustained']['value'], 'e2e', d['e2e']['value'], d['parity']['rel_err'], d['roofline_step'])
for o in json.load(open('gpurun_out/bench_ops.json'))['ops'][24:]: print(f"{o['op']:80s} {o['ms']:.4f} tb {o['t_bound_ms']:.4f} {o['frac']}")
PY
