#!/bin/bash
# compute-sanitizer memcheck over this session's new kernels (fused expansion -> dw -> pw, dw -> pw tail blocks, channel affine)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ref_modules.py -m gpu -q 2>&1 | tail -25
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 120 \
  python -m pytest tests/test_gpu_nets.py tests/test_ref_modules.py -q -x -k "(test_fused_dw_pw and fp16 and not whole) or (blocks_gpu and bf16)" > gpurun_out/r02_sanitize.log 2>&1
echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r02_sanitize.log | head -20
