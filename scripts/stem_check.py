#!/usr/bin/env python
"""Isolated parity + op-name check of the space-to-depth stem kernel (conv_igemm3s.cu) through the mirror blocks.
Each case runs in a child process with a timeout so a hung mbarrier pipeline cannot take the box down."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [  # name, k, Cout, act, N, H, W
    ("stem7_64_small", 7, 64, "relu", 3, 64, 64),
    ("stem7_64_224", 7, 64, "relu", 2, 224, 224),
    ("stem3_32_relu6_224", 3, 32, "relu6", 2, 224, 224),
    ("stem3_64_480", 3, 64, "relu", 1, 480, 480),
    ("stem3_32_ragged", 3, 32, "relu", 3, 50, 38),
    ("stem7_64_ragged", 7, 64, "relu", 5, 38, 50),
    ("stem5_64", 5, 64, "relu", 2, 96, 96),
    # conv -> MaxPool2d(3, 2, 1) fused into the stem kernel (PCV_CONV_POOL3S2)
    ("pool_stem7_64_224", 7, 64, "relu", 3, 224, 224),
    ("pool_stem7_64_bs40", 7, 64, "relu", 40, 224, 224),
    ("pool_stem7_64_small", 7, 64, "relu", 5, 64, 64),
    ("pool_stem3_64_160", 3, 64, "relu", 2, 160, 96),
    ("pool_stem7_64_448", 7, 64, "relu", 1, 448, 448),
    # fused-pool geometry edges: 124-column map (127-pixel s2d rows), R = 2 (Ho % 4 != 0), 80 columns, 3x3 stem, ReLU6
    ("pool_stem7_64_w248", 7, 64, "relu", 2, 200, 248),
    ("pool_stem7_64_h228", 7, 64, "relu", 3, 228, 224),
    ("pool_stem7_64_w160", 7, 64, "relu", 5, 256, 160),
    ("pool_stem3_64_224", 3, 64, "relu6", 2, 224, 224),
    ("pool_stem5_64_192", 5, 64, "relu", 2, 192, 192),
    # fp16 storage tier (the f16 build of the same kernel; name prefix selects the tier)
    ("f16_stem7_64_224", 7, 64, "relu", 2, 224, 224),
    ("f16_stem3_32_relu6_224", 3, 32, "relu6", 2, 224, 224),
    ("f16_pool_stem7_64_224", 7, 64, "relu", 3, 224, 224),
    ("f16_pool_stem3_64_160", 3, 64, "relu", 2, 160, 96),
]


def run(idx):
    import torch
    import torch.nn.functional as F
    import pytorchcv_b200 as P
    from pytorchcv_b200 import blocks
    from oracle import seeded_init, seeded_input
    name, k, cout, act, N, H, W = CASES[idx]
    blk = blocks.ConvBlock(in_channels=3, out_channels=cout, kernel_size=k, stride=2, padding=k // 2,
                           activation=(lambda: torch.nn.ReLU6(inplace=True)) if act == "relu6" else (lambda: torch.nn.ReLU(inplace=True)))
    half = name.startswith("f16_")
    tdt, tol = (torch.float16, 2e-3) if half else (torch.bfloat16, 1.2e-2)
    pool = "pool_" in name
    if pool:
        blk = torch.nn.Sequential(blk, torch.nn.MaxPool2d(kernel_size=3, stride=2, padding=1))
    blk = seeded_init(blk.eval(), seed=3, randomize_bn=True)
    x = seeded_input((N, 3, H, W), seed=5)
    cb = blk[0] if pool else blk
    bn = cb.bn
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    rnd = lambda t: t.to(tdt).float()
    wf = rnd(cb.conv.weight * scale.view(-1, 1, 1, 1))
    bf = bn.bias - bn.running_mean * scale
    with torch.no_grad():
        ref = F.conv2d(rnd(x), wf, bf, stride=2, padding=k // 2)
        ref = ref.clamp(0, 6) if act == "relu6" else torch.relu(ref)
        if pool:
            ref = F.max_pool2d(ref, 3, 2, 1)
    fast = P.accelerate(blk.cuda(), dtype="fp16" if half else "bf16", graph=False)
    y = fast(x.cuda()).float().cpu()
    torch.cuda.synchronize()
    cm = fast.compiled(x.cuda())
    names = [r[0] for r in cm.profile()]
    rel = float((y - ref).abs().max() / ref.abs().max())
    out = {"case": name, "rel": rel, "ok": bool(rel <= tol and torch.isfinite(y).all()), "ops": names}
    if not out["ok"]:
        err = (y - ref).abs() > tol * ref.abs().max()
        out["bad_frac"] = float(err.float().mean())
        out["bad_by_channel"] = [round(v, 2) for v in err.float().mean(dim=(0, 2, 3))[:16].tolist()]
        out["bad_by_row"] = [round(v, 2) for v in err.float().mean(dim=(0, 1, 3))[:16].tolist()]
        out["bad_by_col"] = [round(v, 2) for v in err.float().mean(dim=(0, 1, 2))[:16].tolist()]
        out["got"] = [round(v, 3) for v in y[0, :6, 0, 0].tolist()]
        out["ref"] = [round(v, 3) for v in ref[0, :6, 0, 0].tolist()]
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1:
        print(json.dumps(run(int(sys.argv[1]))), flush=True)
        sys.exit(0)
    fails = 0
    for i, c in enumerate(CASES):
        try:
            r = subprocess.run([sys.executable, "-u", __file__, str(i)], capture_output=True, text=True, timeout=150)
            line = (r.stdout.strip().splitlines() or ["{}"])[-1]
            print(c[0], "rc", r.returncode, line[:900], flush=True)
            if r.returncode != 0:
                print(r.stderr[-800:], flush=True)
                fails += 1
            elif not json.loads(line).get("ok"):
                fails += 1
        except subprocess.TimeoutExpired:
            print(c[0], "TIMEOUT", flush=True)
            fails += 1
    print("stem_check fails:", fails)
