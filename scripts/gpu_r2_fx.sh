#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nets.py -q -x -k "fused_bottleneck" 2>&1 | tail -25 > gpurun_out/r02_fx_test.log; grep -E "passed|failed|Error|error|assert" gpurun_out/r02_fx_test.log | head -12
timeout 600 python -m pytest tests/test_gpu_nets.py -q -k "resnet50 and (bf16 or fp16 or benchmarked)" 2>&1 | tail -8
timeout 600 python bench.py --steps 50 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/r02_fx_bench.json 2> gpurun_out/r02_fx_bench.err; tail -c 300 gpurun_out/r02_fx_bench.err
PCV_FUSE_TAIL=0 timeout 600 python bench.py --steps 50 --warmup 5 --no-configs --no-cpu-baseline --ops-out gpurun_out/bench_ops_nofuse.json > gpurun_out/r02_fx_bench_nofuse.json 2>/dev/null
python - <<'PY'
import json
for n in ("r02_fx_bench","r02_fx_bench_nofuse"):
    d=json.loads(open(f'gpurun_out/{n}.json').read().strip().splitlines()[-1])
    print(n, 'VALUE', d['value'], d['ms_per_step'], 'sustained', d['sustained']['value'], 'e2e', d['e2e']['value'], d['parity']['rel_err'], d['parity']['top1_equal'], d['roofline_step'])
for o in json.load(open('gpurun_out/bench_ops.json'))['ops'][:14]: print(f"{o['op']:90s} {o['ms']:.4f} tb {o['t_bound_ms']:.4f} {o['frac']}")
PY
