"""Timing experiments on the fused dw->pw kernel (PCV_DP_DBG bits) - per-op ms of one LinearBottleneck."""
import copy, os, sys, torch
import pytorchcv_b200 as P
from pytorchcv_b200 import nets as M, blocks as BK

CASES = [(24, 24, 1, 56, 56), (32, 16, 1, 112, 112), (16, 24, 2, 112, 112), (64, 64, 1, 14, 14)]
for cin, cout, s, h, w in CASES:
    exp = not (cin == 32 and cout == 16)
    unit = M.LinearBottleneck(cin, cout, s, expansion=exp, remove_exp_conv=True, activation=BK.lambda_relu6()).eval().cuda()
    x = torch.randn(256, cin, h, w, device="cuda")
    fast = P.accelerate(unit, dtype="fp16", graph=False)
    fast(x)
    c = fast.compiled(x)
    best = {}
    for _ in range(5):
        for nm, ms, *_ in c.profile():
            if "dwpw" in nm or "dwconv" in nm:
                best[nm] = min(best.get(nm, 1e9), ms)
    for k, v in best.items():
        print("DBG=%s  %-80s %.4f" % (os.environ.get("PCV_DP_DBG", "0"), k, v), flush=True)
