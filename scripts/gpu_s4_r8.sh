#!/bin/bash
mkdir -p gpurun_out
run() {  # name, lib, dbg
  if [ -n "$2" ]; then export PCV_B200_LIB=$PWD/pytorchcv_b200/$2; else unset PCV_B200_LIB; fi
  PCV_STEM_DBG=$3 timeout 300 python bench.py --no-cpu-baseline --steps 30 --ops-out gpurun_out/abc_ops_$1.json > gpurun_out/abc_$1.json 2> gpurun_out/abc_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/abc_$1.json").read().strip().splitlines()[-1]); o=json.load(open("gpurun_out/abc_ops_$1.json")); print("$1", d["value"], d["ms_per_step"], "stem", o["ops"][0]["ms"], "sum", round(sum(r["ms"] for r in o["ops"]),4))
except Exception as e: print("$1 failed", e); print(open("gpurun_out/abc_$1.err").read()[-800:])
PY
}
for rep in 1 2; do
run prev libpcv_b200_prev.so 0
run mid0 libpcv_b200_mid.so 0
run mid16 libpcv_b200_mid.so 16
run mid32 libpcv_b200_mid.so 32
run mid48 libpcv_b200_mid.so 48
run new0 "" 0
run new48 "" 48
done
