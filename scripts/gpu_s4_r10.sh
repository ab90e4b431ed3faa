#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/gpu_diag.py --timeout 60 --only halo --out gpurun_out/diag_halo.jsonl 2>&1 | cut -c1-400 | tail -9
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log | cut -c1-600
bash scripts/gpu_ab.sh
