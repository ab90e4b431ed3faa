"""Run one fused conv layer in isolation (for ncu captures and quick timing).

    python scripts/profile_layer.py --N 256 --H 56 --Cin 64 --Cout 256 --k 1 [--res 1] [--reps 5]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from pytorchcv_b200 import functional as P, _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    for n, d in (("N", 256), ("H", 56), ("Cin", 64), ("Cout", 256), ("k", 1), ("stride", 1), ("dil", 1), ("groups", 1),
                 ("res", 0), ("act", 1), ("reps", 5), ("flags", 0)):
        ap.add_argument("--" + n, type=int, default=d)
    a = ap.parse_args()
    pad = a.dil * (a.k // 2)
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(0)
    x = torch.randn(a.N, a.H, a.H, a.Cin, generator=g).to(dev).to(torch.bfloat16)
    w = (torch.randn(a.Cout, a.Cin // a.groups, a.k, a.k, generator=g) * 0.05).to(dev)
    desc = P.make_desc(a.N, a.H, a.H, a.Cin, a.Cout, a.k, a.stride, pad, a.dil, a.groups, a.act, flags=a.flags)
    packed = P.pack_conv(desc, _lib.BF16, w)
    Ho = (a.H + 2 * pad - a.dil * (a.k - 1) - 1) // a.stride + 1
    res = torch.randn(a.N, Ho, Ho, a.Cout, generator=g).to(dev).to(torch.bfloat16) if a.res else None
    out = torch.empty(a.N, Ho, Ho, a.Cout, dtype=torch.bfloat16, device=dev)
    for _ in range(2):
        P.conv2d(x, packed, res, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        P.conv2d(x, packed, res, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    M = a.N * Ho * Ho
    fl = 2.0 * M * a.Cout * (a.Cin // a.groups) * a.k * a.k
    by = 2.0 * (a.N * a.H * a.H * a.Cin + M * a.Cout * (2 if a.res else 1)) + 2.0 * w.numel()
    print(f"layer k{a.k} s{a.stride} {a.Cin}->{a.Cout} @{a.H} N={a.N} res={a.res}: {ms:.4f} ms  "
          f"{fl / ms / 1e9:.1f} TFLOP/s  {by / ms / 1e6:.1f} GB/s")


if __name__ == "__main__":
    main()
