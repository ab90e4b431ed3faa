#!/bin/bash
# ncu --set full on isolated hot-path ops (scripts/profile_ops.py); one .ncu-rep back in gpurun_out/
mkdir -p gpurun_out
ONLY=${1:-}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'igemm|maxpool|dwconv|gap_kernel' \
   -o gpurun_out/ops_full -f python scripts/profile_ops.py --set ${SET:-all} --reps 1 --warm 0 ${ONLY:+--only $ONLY} > gpurun_out/ncu_ops.log 2>&1
tail -5 gpurun_out/ncu_ops.log
ls -la gpurun_out/*.ncu-rep
