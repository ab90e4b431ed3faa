#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_nets.py -q -k "fp32 or f32" 2>&1 | tail -30 > gpurun_out/r02_f32_nets.log; grep -E "passed|failed|FAILED|AssertionError: \(|assert 0\." gpurun_out/r02_f32_nets.log | head -20
