#!/bin/bash
for dbg in 16 19 20 23; do
echo "== dbg $dbg"
PCV_F3_DBG=$dbg timeout 300 python - <<'PY' 2>&1 | sed -n "/=== timed pass/,\$p" | sed -n 4,5p
import torch, sys
sys.path.insert(0, '.')
import pytorchcv_b200 as P
from bench import build_net
net = build_net("resnet18", 224, 224).cuda()
fast = P.accelerate(net, dtype="fp32", graph=False)
x = torch.randn(8, 3, 224, 224, device="cuda")
for _ in range(3):
    fast(x)
torch.cuda.synchronize()
print("=== timed pass", flush=True)
fast(x)
torch.cuda.synchronize()
PY
done
