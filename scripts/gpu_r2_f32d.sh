#!/bin/bash
echo "== SIMT chain"; PCV_F32_SPLIT=0 timeout 600 python scripts/f32_unit_diag.py seresnext50_32x4d 1 chain 2>&1 | awk '{print $(NF-2)}' | tr '\n' ' '; echo
for ch in 1 4 16; do echo "== chunk $ch chain"; PCV_F3_CHUNK=$ch timeout 600 python scripts/f32_unit_diag.py seresnext50_32x4d 1 chain 2>&1 | awk '{print $(NF-2)}' | tr '\n' ' '; echo; done
