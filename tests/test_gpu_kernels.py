"""Per-kernel parity on a real B200: every C-ABI op against the CPU oracle's leaf ops (torch CPU fp32 functional),
on identical seeded inputs.  Tolerances: fp32 tier 1e-4 (north star), bf16 tier 1.2e-2 per fused block with the
oracle fed bf16-rounded operands (so the bound checks the kernel, not the storage format)."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "scripts"))
import gpu_diag  # noqa: E402  (shares the conv case table with the first-light diagnostics)
import stem_check  # noqa: E402  (space-to-depth stem cases, incl. the fused MaxPool2d(3, 2, 1))

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.mark.parametrize("idx", range(len(gpu_diag.CASES)), ids=[c[0] for c in gpu_diag.CASES])
def test_conv_block_kernels(idx):
    r = gpu_diag.run_case(idx)
    assert r["ok"], r


@pytest.mark.parametrize("idx", range(len(stem_check.CASES)), ids=[c[0] for c in stem_check.CASES])
def test_s2d_stem_kernels(idx):
    """ConvBlock(k x k, stride 2) on a 3-channel image [-> MaxPool2d(3, 2, 1)] (resnet.py:232-263, mobilenetv2.py:101,
    senet.py:127-164) through the space-to-depth halo kernel; pooled cases must report the fused op."""
    r = stem_check.run(idx)
    assert r["ok"], r
    assert r["ops"][0].startswith("conv_stem"), r["ops"]
    name, k, cout, _, _, H, W = stem_check.CASES[idx]
    if "pool_" in name:
        from pytorchcv_b200 import _lib
        fused = bool(_lib.load().pcv_stem_s2d_pool_ok(3, H, W, k, cout))   # conv maps 64..125 columns wide fuse the pool
        assert ("+maxpool3s2" in r["ops"][0]) == fused, r["ops"]
        assert len(r["ops"]) == (2 if fused else 3), r["ops"]
        assert fused == (64 <= W // 2 <= 128 - k // 2)


def _nhwc(x, dtype):
    return x.cuda().permute(0, 2, 3, 1).contiguous().to(dtype)


def _nchw(y):
    return y.float().cpu().permute(0, 3, 1, 2)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("shape", [(2, 64, 112, 112), (1, 128, 15, 15), (3, 8, 7, 9)])
def test_maxpool_3x3_s2_p1(shape, dtype):
    from pytorchcv_b200 import functional as P
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(1)) - 1.5  # mostly negative: -inf padding matters
    xr = x.to(dtype).float()
    want = F.max_pool2d(xr, 3, 2, 1)
    got = _nchw(P.maxpool2d(_nhwc(x, dtype), 3, 2, 1))
    assert torch.equal(got, want)  # max of representable values: exact


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 4e-3), (torch.float16, 1e-3), (torch.float32, 1e-6)])
@pytest.mark.parametrize("shape", [(4, 2048, 7, 7), (2, 256, 56, 56), (1, 72, 5, 3)])
def test_global_avgpool(shape, dtype, tol):
    from pytorchcv_b200 import functional as P
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(2)) + 0.3
    want = x.to(dtype).float().mean(dim=(2, 3))
    got32 = P.global_avgpool(_nhwc(x, dtype), out_dtype=torch.float32).cpu()
    assert _rel(got32, want) <= 1e-5
    got = P.global_avgpool(_nhwc(x, dtype)).float().cpu()
    assert _rel(got, want) <= tol


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 4e-3), (torch.float16, 1e-3), (torch.float32, 1e-6)])
@pytest.mark.parametrize("shape,k", [((2, 512, 60, 60), 6), ((1, 264, 15, 17), 3), ((3, 64, 7, 7), 2), ((2, 2048, 9, 9), 1)])
def test_adaptive_avgpool(shape, k, dtype, tol):
    """nn.AdaptiveAvgPool2d(k) of PyramidPoolingBranch (pspnet.py:71): torch's floor/ceil bin edges, overlapping bins."""
    from pytorchcv_b200 import functional as P
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(13))
    want = F.adaptive_avg_pool2d(x.to(dtype).float(), k)
    got = _nchw(P.adaptive_avgpool(_nhwc(x, dtype), k, k))
    assert got.shape == want.shape
    assert _rel(got, want) <= tol


@pytest.mark.parametrize("N,C,mid", [(4, 256, 16), (3, 2048, 128), (1, 64, 4), (70, 512, 32), (256, 1000, 60)])
def test_se_excite_and_scale(N, C, mid):
    from pytorchcv_b200 import functional as P, _lib
    g = torch.Generator().manual_seed(3)
    pooled = torch.randn(N, C, generator=g)
    w1, b1 = torch.randn(mid, C, generator=g) * 0.1, torch.randn(mid, generator=g) * 0.1
    w2, b2 = torch.randn(C, mid, generator=g) * 0.3, torch.randn(C, generator=g) * 0.1
    want = torch.sigmoid(F.linear(torch.relu(F.linear(pooled, w1, b1)), w2, b2))
    gate = P.se_excite(pooled.cuda(), w1.cuda(), b1.cuda(), w2.cuda(), b2.cuda())
    assert _rel(gate.cpu(), want) <= 1e-5
    x = torch.randn(N, C, 6, 5, generator=g)
    idn = torch.randn(N, C, 6, 5, generator=g)
    for dtype, tol in ((torch.float32, 1e-6), (torch.bfloat16, 8e-3), (torch.float16, 1e-3)):
        xr, ir = x.to(dtype).float(), idn.to(dtype).float()
        ref = torch.relu(xr * want[:, :, None, None] + ir)
        got = _nchw(P.se_scale_add_act(_nhwc(x, dtype), gate, _nhwc(idn, dtype), _lib.ACT_RELU))
        assert _rel(got, ref) <= max(tol, 1e-5)
        ref2 = xr * want[:, :, None, None]
        got2 = _nchw(P.se_scale_add_act(_nhwc(x, dtype), gate, None, _lib.ACT_NONE))
        assert _rel(got2, ref2) <= max(tol, 1e-5)


def test_add_act_and_layout_roundtrip():
    from pytorchcv_b200 import functional as P, _lib
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 3, 17, 13, generator=g)
    nhwc = P.nchw_to_nhwc(x.cuda(), torch.float32)                    # channels padded 3 -> 8 with zeros
    assert nhwc.shape == (2, 17, 13, 8) and float(nhwc[..., 3:].abs().max()) == 0.0
    assert torch.equal(P.nhwc_to_nchw(nhwc, channels=3).cpu(), x)     # fp32 round trip is exact
    for dt in (torch.bfloat16, torch.float16):
        back = P.nhwc_to_nchw(P.nchw_to_nhwc(x.cuda(), dt), channels=3).cpu()
        assert torch.equal(back, x.to(dt).float())                    # 16-bit round trip == one rounding
    a, b = torch.randn(2, 5, 5, 16, generator=g), torch.randn(2, 5, 5, 16, generator=g)
    got = P.add_act(a.cuda(), b.cuda(), _lib.ACT_RELU6).cpu()
    assert torch.equal(got, (a + b).clamp(0, 6))


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 8e-3), (torch.float16, 1e-3)])
def test_bilinear_align_corners(dtype, tol):
    from pytorchcv_b200 import functional as P
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 21, 15, 15, generator=g)
    want = F.interpolate(x.to(dtype).float(), size=(120, 120), mode="bilinear", align_corners=True)
    xin = torch.zeros(2, 15, 15, 24, dtype=dtype, device="cuda")      # 21 logical channels at pitch 24
    xin[..., :21] = _nhwc(x, dtype)
    got = P.bilinear_upsample_ac(xin, 120, 120, channels=21, nchw_f32=True).cpu()
    assert _rel(got, want) <= 1e-5   # fp32 output: the source rounding is in `want`, the lerp itself is fp32
    for size in ((24, 24), (44, 32)):                                 # ratios < 2 and 2..3: the per-element / 3-pixel paths
        want_s = F.interpolate(x.to(dtype).float(), size=size, mode="bilinear", align_corners=True)
        got_s = P.bilinear_upsample_ac(xin, size[0], size[1], channels=21, nchw_f32=True).cpu()
        assert _rel(got_s, want_s) <= 1e-5
    x2 = torch.randn(2, 16, 1, 1, generator=g)                        # 1x1 source == pure broadcast (ASPP avg branch)
    got2 = _nchw(P.bilinear_upsample_ac(_nhwc(x2, dtype), 9, 9))
    assert _rel(got2, x2.to(dtype).float().expand(2, 16, 9, 9)) <= tol


def test_plan_records_and_replays():
    """The same op recorded into a plan replays bit-identically, also from a CUDA graph."""
    import ctypes as C
    from pytorchcv_b200 import functional as P, _lib
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 64, 14, 14, generator=g)
    w = torch.randn(64, 64, 3, 3, generator=g) * 0.05
    desc = P.make_desc(2, 14, 14, 64, 64, 3, 1, 1, 1, 1, _lib.ACT_RELU)
    packed = P.pack_conv(desc, _lib.BF16, w.cuda())
    xin = _nhwc(x, torch.bfloat16)
    eager = P.conv2d(xin, packed).clone()
    out = torch.zeros_like(eager)
    plan = C.c_void_p()
    _lib.call("pcv_plan_create", C.byref(plan))
    _lib.call("pcv_conv2d_bias_act", plan, C.byref(desc), _lib.BF16, xin.data_ptr(), packed.w.data_ptr(),
              packed.bias.data_ptr(), None, out.data_ptr(), None)
    assert _lib.load().pcv_plan_num_ops(plan) == 1
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        _lib.call("pcv_plan_run", plan, s.cuda_stream)
        s.synchronize()
        assert torch.equal(out, eager)
        out.zero_()
        _lib.call("pcv_plan_graph_launch", plan, s.cuda_stream)
        _lib.call("pcv_plan_graph_launch", plan, s.cuda_stream)
        s.synchronize()
        assert torch.equal(out, eager)
    name = _lib.load().pcv_plan_op_name(plan, 0).decode()
    assert name.startswith("conv_tc") and " 3x3 " in name
    _lib.call("pcv_plan_destroy", plan)
