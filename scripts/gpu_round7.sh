#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gpu_diag.py --only dw --timeout 40 --out gpurun_out/diag_dw.jsonl 2>&1 | cut -c1-300
timeout 100 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "maxpool" 2>&1 | tail -3
timeout 120 python scripts/profile_ops.py --set mobilenet,pool > gpurun_out/profile_ops5.log 2>&1; cat gpurun_out/profile_ops5.log
