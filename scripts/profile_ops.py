"""Run a fixed list of hot-path ops in isolation, one process, for `ncu` captures and quick CUDA-event timing.

    python scripts/profile_ops.py [--set resnet50|mobilenet|se|all] [--reps 3] [--only SUBSTR[,SUBSTR...]]

Each op is launched `--warm` times untimed and `--reps` times timed; L2 is flushed (a 256 MB memset) before every
timed launch so the number is the cold-L2 figure the roofline (HBM) is quoted against.  Prints one line per op:
name, ms, TFLOP/s, GB/s (algorithmic bytes of SURVEY 8d).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from pytorchcv_b200 import functional as P, _lib  # noqa: E402

# kind, name, N, H, Cin, Cout, k, stride, dil, groups, res, act
CONVS = {
    "resnet50": [
        ("c3_64_56", 256, 56, 64, 64, 3, 1, 1, 1, 0, 1),
        ("c1_64_256_56_res", 256, 56, 64, 256, 1, 1, 1, 1, 1, 1),
        ("c1_256_64_56", 256, 56, 256, 64, 1, 1, 1, 1, 0, 1),
        ("c3_128_28", 256, 28, 128, 128, 3, 1, 1, 1, 0, 1),
        ("c1_128_512_28_res", 256, 28, 128, 512, 1, 1, 1, 1, 1, 1),
        ("c3_256_14", 256, 14, 256, 256, 3, 1, 1, 1, 0, 1),
        ("c1_256_1024_14_res", 256, 14, 256, 1024, 1, 1, 1, 1, 1, 1),
        ("c1_1024_256_14", 256, 14, 1024, 256, 1, 1, 1, 1, 0, 1),
        ("c3_512_7", 256, 7, 512, 512, 3, 1, 1, 1, 0, 1),
        ("c1_512_2048_7_res", 256, 7, 512, 2048, 1, 1, 1, 1, 1, 1),
        ("c1_2048_512_7", 256, 7, 2048, 512, 1, 1, 1, 1, 0, 1),
        ("c1s2_512_1024_28", 256, 28, 512, 1024, 1, 2, 1, 1, 0, 0),
    ],
    "se": [
        ("g3_128_56", 256, 56, 128, 128, 3, 1, 1, 32, 0, 1),
        ("g3_256_28", 256, 28, 256, 256, 3, 1, 1, 32, 0, 1),
        ("g3_512_14", 256, 14, 512, 512, 3, 1, 1, 32, 0, 1),
        ("g3_1024_7", 256, 7, 1024, 1024, 3, 1, 1, 32, 0, 1),
    ],
    "mobilenet": [
        ("dw3_32_112", 256, 112, 32, 32, 3, 1, 1, 32, 0, 2),
        ("dw3s2_96_112", 256, 112, 96, 96, 3, 2, 1, 96, 0, 2),
        ("dw3_144_56", 256, 56, 144, 144, 3, 1, 1, 144, 0, 2),
        ("dw3_192_28", 256, 28, 192, 192, 3, 1, 1, 192, 0, 2),
        ("dw3_384_14", 256, 14, 384, 384, 3, 1, 1, 384, 0, 2),
        ("dw3_960_7", 256, 7, 960, 960, 3, 1, 1, 960, 0, 2),
        ("c1_16_96_112", 256, 112, 16, 96, 1, 1, 1, 1, 0, 2),
        ("c1_144_24_56_res", 256, 56, 144, 24, 1, 1, 1, 1, 1, 0),
    ],
    "deeplab": [   # DeepLabv3 bs16 480x480: ASPP dilated 3x3 (the dominant kernel), backbone dilated 3x3
        ("c3_d12_2048_256_60", 16, 60, 2048, 256, 3, 1, 12, 1, 0, 1),
        ("c3_d2_256_256_60", 16, 60, 256, 256, 3, 1, 2, 1, 0, 1),
    ],
    "effi": [   # EfficientNet-b0 / MobileNetV3: 5x5 depthwise, swish epilogues (act 4)
        ("dw5_240_28_swish", 256, 28, 240, 240, 5, 1, 1, 240, 0, 4),
        ("dw5s2_144_56_swish", 256, 56, 144, 144, 5, 2, 1, 144, 0, 4),
        ("dw5_672_14_swish", 256, 14, 672, 672, 5, 1, 1, 672, 0, 4),
        ("c1_16_96_112_swish", 256, 112, 16, 96, 1, 1, 1, 1, 0, 4),
    ],
}


def flush(buf):
    buf.zero_()


def time_op(fn, reps, warm, buf):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush(buf)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / max(reps, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--set", default="all")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--warm", type=int, default=1)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(0)
    buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sets = list(CONVS) if a.set == "all" else a.set.split(",")
    for s in sets:
        for (name, N, H, Cin, Cout, k, stride, dil, groups, res, act) in CONVS.get(s, []):
            if a.only and not any(o in name for o in a.only.split(",")):
                continue
            pad = dil * (k // 2)
            x = torch.randn(N, H, H, Cin, generator=g).to(dev).to(torch.bfloat16)
            w = (torch.randn(Cout, Cin // groups, k, k, generator=g) * 0.05).to(dev)
            desc = P.make_desc(N, H, H, Cin, Cout, k, stride, pad, dil, groups, act)
            packed = P.pack_conv(desc, _lib.BF16, w)
            Ho = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
            r = torch.randn(N, Ho, Ho, Cout, generator=g).to(dev).to(torch.bfloat16) if res else None
            out = torch.empty(N, Ho, Ho, Cout, dtype=torch.bfloat16, device=dev)
            ms = time_op(lambda: P.conv2d(x, packed, r, out), a.reps, a.warm, buf)
            M = N * Ho * Ho
            fl = 2.0 * M * Cout * (Cin // groups) * k * k
            pin = Ho * Ho if (k == 1 and stride > 1) else H * H
            by = 2.0 * (N * pin * Cin + M * Cout * (2 if res else 1)) + 2.0 * w.numel()
            print(f"{name:22s} {ms:8.4f} ms  {fl / ms / 1e9:8.1f} TFLOP/s  {by / ms / 1e6:8.1f} GB/s", flush=True)
            del x, w, packed, r, out
    if a.set in ("all", "pool") or "pool" in sets:
        for (name, N, H, C) in (("maxpool_64_112", 256, 112, 64), ("maxpool_128_240", 16, 240, 128)):
            if a.only and not any(o in name for o in a.only.split(",")):
                continue
            x = torch.randn(N, H, H, C, generator=g).to(dev).to(torch.bfloat16)
            ms = time_op(lambda: P.maxpool2d(x, 3, 2, 1), a.reps, a.warm, buf)
            by = 2.0 * N * C * (H * H + (H // 2) ** 2)
            print(f"{name:22s} {ms:8.4f} ms  {0.0:8.1f} TFLOP/s  {by / ms / 1e6:8.1f} GB/s", flush=True)
        for (name, N, HW, C) in (("sescale_256_3136", 256, 56, 256),):
            if a.only and not any(o in name for o in a.only.split(",")):
                continue
            x = torch.randn(N, HW, HW, C, generator=g).to(dev).to(torch.bfloat16)
            idn = torch.randn(N, HW, HW, C, generator=g).to(dev).to(torch.bfloat16)
            gate = torch.rand(N, C, generator=g).to(dev)
            ms = time_op(lambda: P.se_scale_add_act(x, gate, idn, _lib.ACT_RELU), a.reps, a.warm, buf)
            by = 2.0 * N * C * HW * HW * 3
            print(f"{name:22s} {ms:8.4f} ms  {0.0:8.1f} TFLOP/s  {by / ms / 1e6:8.1f} GB/s", flush=True)
        for (name, N, HW, C) in (("gavg_2048_49", 256, 7, 2048), ("gavg_256_3136", 256, 56, 256)):
            if a.only and not any(o in name for o in a.only.split(",")):
                continue
            x = torch.randn(N, HW, HW, C, generator=g).to(dev).to(torch.bfloat16)
            ms = time_op(lambda: P.global_avgpool(x, out_dtype=torch.float32), a.reps, a.warm, buf)
            by = 2.0 * N * C * HW * HW
            print(f"{name:22s} {ms:8.4f} ms  {0.0:8.1f} TFLOP/s  {by / ms / 1e6:8.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
