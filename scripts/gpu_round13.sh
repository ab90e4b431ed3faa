#!/bin/bash
mkdir -p gpurun_out
for cfg in "0 0 0" "2 2 6" "3 1 8" "4 1 6" "5 1 4" "2 1 8" "6 1 2" "2 3 2"; do
set -- $cfg
if [ "$1" = "0" ]; then unset PCV_IGEMM2_STAGES PCV_IGEMM2_KSUB PCV_IGEMM2_NSTG; else export PCV_IGEMM2_STAGES=$1 PCV_IGEMM2_KSUB=$2 PCV_IGEMM2_NSTG=$3; fi
echo "== stages=$1 ksub=$2 nstg=$3"
timeout 100 python scripts/profile_ops.py --set resnet50 --reps 3 2>&1 | awk '{printf "%s %s | ", $1, $2} END {print ""}'
done
