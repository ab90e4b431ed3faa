#!/bin/bash
# compute-sanitizer memcheck over this round's new kernels / code paths: fused expansion -> dw -> pw, dw -> pw tail blocks,
# channel affine pass, gated conv3 epilogue, N = 16 grouped MMAs, virtual-channel-padding lowerings (GhostNet / MixNet blocks)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 120 \
  python -m pytest tests/test_gpu_nets.py tests/test_ref_modules.py tests/test_gpu_kernels.py -q -x \
  -k "(test_fused_dw_pw and fp16 and not whole) or (blocks_gpu and bf16) or se_gate_in_conv3 or grouped" > gpurun_out/r02_sanitize.log 2>&1
echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r02_sanitize.log | head -20
