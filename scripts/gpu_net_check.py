"""Whole-network parity check on a real B200: compiled plan (bf16 / fp32 tiers) vs the CPU oracle and the golden
fixtures generated from the reference.  Each (net, tier) runs in a subprocess with a timeout.

    python scripts/gpu_net_check.py [--only resnet] [--timeout 300]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NETS = [
    ("resnet18", (2, 3, 224, 224), "resnet18_bs2"),
    ("resnet50", (2, 3, 224, 224), "resnet50_bs2"),
    ("mobilenetv2_w1", (2, 3, 224, 224), "mobilenetv2_w1_bs2"),
    ("seresnext50_32x4d", (2, 3, 224, 224), "seresnext50_32x4d_bs2"),
    ("mobilenet_w1", (2, 3, 224, 224), "mobilenet_w1_bs2"),
    ("deeplabv3_resnetd50b_voc", (1, 3, 480, 480), "deeplabv3_resnetd50b_voc_bs1"),
]


def run(idx: int, tier: str) -> dict:
    import numpy as np
    import torch
    import pytorchcv_b200 as P
    from oracle import oracle_forward, seeded_init, seeded_input

    name, shape, stem = NETS[idx]
    net = seeded_init(P.get_model(name, pretrained=False).eval(), seed=0, randomize_bn=True)
    x = seeded_input(shape, seed=1234)
    ref = oracle_forward(net, x)
    refs = ref if isinstance(ref, (tuple, list)) else (ref,)
    gold = np.load(os.path.join(ROOT, "tests", "golden", stem + ".npz"))
    fast = P.accelerate(net.cuda(), dtype=tier)
    y = fast(x.cuda())
    torch.cuda.synchronize()
    ys = y if isinstance(y, (tuple, list)) else (y,)
    out = {"net": name, "tier": tier, "ops": fast.compiled(x.cuda()).num_ops}
    rels, agree, grel = [], [], []
    for i, (a, b) in enumerate(zip(ys, refs)):
        a = a.float().cpu()
        rels.append(((a - b).abs().max() / b.abs().max()).item())
        agree.append((a.argmax(1) == b.argmax(1)).float().mean().item())
        g = torch.from_numpy(gold[f"out{i}"])
        sub = 16 if a.dim() == 4 else 1
        asub = a[..., ::sub, ::sub] if a.dim() == 4 else a
        grel.append(((asub - g).abs().max() / g.abs().max()).item())
    out.update(rel_vs_oracle=rels, argmax_agree=agree, rel_vs_golden=grel,
               finite=bool(all(torch.isfinite(t.float()).all().item() for t in ys)))
    tol = 1e-4 if tier == "fp32" else 2e-2
    out["ok"] = out["finite"] and max(rels) <= tol
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", type=int)
    ap.add_argument("--tier", type=str, default="bf16")
    ap.add_argument("--only", type=str)
    ap.add_argument("--tiers", type=str, default="bf16,fp32")
    ap.add_argument("--timeout", type=int, default=400)
    a = ap.parse_args()
    if a.case is not None:
        print("NETCHK " + json.dumps(run(a.case, a.tier)), flush=True)
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for i, (name, _, _) in enumerate(NETS):
        if a.only and a.only not in name:
            continue
        for tier in a.tiers.split(","):
            try:
                p = subprocess.run([sys.executable, __file__, "--case", str(i), "--tier", tier], capture_output=True,
                                   text=True, timeout=a.timeout)
                line = [ln for ln in p.stdout.splitlines() if ln.startswith("NETCHK ")]
                r = json.loads(line[-1][7:]) if line else {"net": name, "tier": tier, "ok": False,
                                                            "error": (p.stderr or p.stdout)[-800:]}
            except subprocess.TimeoutExpired:
                r = {"net": name, "tier": tier, "ok": False, "error": "TIMEOUT"}
            print(json.dumps(r), flush=True)
            with open(os.path.join(ROOT, "gpurun_out", "netcheck.jsonl"), "a") as f:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
