#!/bin/bash
# what the driver runs at round end: the whole GPU suite, smoke(), the default bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_full_pytest.log; tail -6 gpurun_out/r02_full_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_full_bench.json 2> gpurun_out/r02_full_bench.err
echo "bench rc=$?"; tail -c 500 gpurun_out/r02_full_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_full_bench.json').read().strip().splitlines()[-1])
print('VALUE', d['value'], 'sustained', d.get('sustained',{}).get('value'), 'e2e', d['e2e']['value'], 'e2e_f32', d['e2e_f32']['value'], 'parity', d['parity']['rel_err'], d['parity']['top1_equal'])
print('roof', d['roofline']['kernel'], d['roofline']['frac'], 'step', d['roofline_step'])
print('cpu', d['cpu_baseline'])
for c in d.get('configs', []): print(c.get('model'), c.get('dtype'), c.get('value'), c.get('e2e'), c.get('roofline_step'), (c.get('parity') or {}).get('rel_err'), (c.get('parity') or {}).get('top1_equal'), c.get('error'))
PY
