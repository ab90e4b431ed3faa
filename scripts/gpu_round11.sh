#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 4 5 6 3; do
echo "== PCV_IGEMM_DBG=$d (bit0 no loads, bit1 no MMA, bit2 no epilogue)"
PCV_IGEMM_DBG=$d PCV_IGEMM_HALO=0 timeout 100 python scripts/profile_ops.py --set resnet50 --reps 3 2>&1 | awk '{printf "%s %s | ", $1, $2} END {print ""}'
done
