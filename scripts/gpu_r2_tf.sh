#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_nets.py -q -k "b0b or pad4 or _tf" 2>&1 | tail -15
