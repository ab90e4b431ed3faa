#!/bin/bash
mkdir -p gpurun_out
echo "== dbg=0"; PCV_IGEMM3_DBG=0 timeout 200 python scripts/gpu_diag.py --only halo --timeout 40 --out gpurun_out/diag_halo0.jsonl 2>&1 | cut -c1-900
echo "== dbg=1"; PCV_IGEMM3_DBG=1 timeout 200 python scripts/gpu_diag.py --only halo --timeout 40 --out gpurun_out/diag_halo1.jsonl 2>&1 | cut -c1-900
timeout 120 python scripts/profile_ops.py --set resnet50 --only c3_ > gpurun_out/profile_ops3.log 2>&1; cat gpurun_out/profile_ops3.log
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "se_excite or avgpool or plan" 2>&1 | tail -5
