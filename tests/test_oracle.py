"""Pins the oracle (oracle/ref_forward.py) against outputs of the reference itself.

 - golden replay: fixtures in tests/golden/*.npz were produced by /root/reference on CPU (make_golden.py);
 - live: when /root/reference is importable, the oracle is compared with the reference module's own forward.
Runs on CPU in well under a minute.
"""
import json
import os

import numpy as np
import pytest
import torch

import pytorchcv_b200 as P
from pytorchcv_b200 import blocks as B, nets as M
from oracle import oracle_forward, seeded_init, seeded_input
from conftest import GOLDEN

NETS = [
    ("resnet18_bs2", "resnet18", (2, 3, 224, 224), 1),
    ("resnet50_bs2", "resnet50", (2, 3, 224, 224), 1),
    ("mobilenetv2_w1_bs2", "mobilenetv2_w1", (2, 3, 224, 224), 1),
    ("seresnext50_32x4d_bs2", "seresnext50_32x4d", (2, 3, 224, 224), 1),
    ("mobilenet_w1_bs2", "mobilenet_w1", (2, 3, 224, 224), 1),
    ("deeplabv3_resnetd50b_voc_bs1", "deeplabv3_resnetd50b_voc", (1, 3, 480, 480), 16),
    ("efficientnet_b0_bs2", "efficientnet_b0", (2, 3, 224, 224), 1),
    ("mobilenetv3_large_w1_bs2", "mobilenetv3_large_w1", (2, 3, 224, 224), 1),
    ("mobilenetv3_small_w1_bs2", "mobilenetv3_small_w1", (2, 3, 224, 224), 1),
    ("seresnet18_bs2", "seresnet18", (2, 3, 224, 224), 1),
    ("seresnet50_bs2", "seresnet50", (2, 3, 224, 224), 1),
    ("fcn8sd_resnetd50b_voc_bs1", "fcn8sd_resnetd50b_voc", (1, 3, 480, 480), 16),
    ("pspnet_resnetd50b_voc_bs1", "pspnet_resnetd50b_voc", (1, 3, 480, 480), 16),
    ("mnasnet_a1_bs2", "mnasnet_a1", (2, 3, 224, 224), 1),
    ("mnasnet_small_bs2", "mnasnet_small", (2, 3, 224, 224), 1),
    ("fbnet_cb_bs2", "fbnet_cb", (2, 3, 224, 224), 1),
    ("spnasnet_bs2", "spnasnet", (2, 3, 224, 224), 1),
    ("senet16_bs2", "senet16", (2, 3, 224, 224), 1),
    ("proxylessnas_mobile_bs2", "proxylessnas_mobile", (2, 3, 224, 224), 1),
    ("efficientnet_b0b_bs2", "efficientnet_b0b", (2, 3, 224, 224), 1),           # tf_mode: asymmetric "SAME" padding
]

BLOCKS = {
    "convblock_3x3_s2": (lambda: B.conv3x3_block(in_channels=16, out_channels=24, stride=2), (2, 16, 15, 15)),
    "convblock_1x1_noact": (lambda: B.ConvBlock(32, 64, kernel_size=1, activation=None), (2, 32, 9, 9)),
    "convblock_3x3_d2_bias": (lambda: B.ConvBlock(16, 16, kernel_size=3, padding=2, dilation=2, bias=True),
                              (1, 16, 12, 12)),
    "dws_3x3": (lambda: B.DwsConvBlock(16, 32, kernel_size=3, stride=1, padding=1), (2, 16, 10, 10)),
    "dwconv5x5_relu6": (lambda: B.dwconv5x5_block(in_channels=24, out_channels=24, activation=B.lambda_relu6()),
                        (1, 24, 11, 11)),
    "seblock_64": (lambda: B.SEBlock(channels=64), (2, 64, 7, 7)),
    "resunit_bottleneck_s2": (lambda: M.ResUnit(64, 128, stride=2, bottleneck=True, conv1_stride=True),
                              (2, 64, 14, 14)),
    "resunit_basic": (lambda: M.ResUnit(32, 32, stride=1, bottleneck=False), (2, 32, 8, 8)),
    "linear_bottleneck_res": (lambda: M.LinearBottleneck(24, 24, stride=1, expansion=True, remove_exp_conv=False,
                                                         activation=B.lambda_relu6()), (2, 24, 14, 14)),
    "seresnext_unit": (lambda: M.SEResNeXtUnit(256, 256, stride=1, cardinality=32, bottleneck_width=4),
                       (1, 256, 8, 8)),
    "effi_dws_unit": (lambda: M.EffiDwsConvUnit(32, 16, stride=1, normalization=B.lambda_batchnorm2d(),
                                                activation=B.lambda_swish(), tf_mode=False), (2, 32, 16, 16)),
    "effi_invres_k5_se": (lambda: M.EffiInvResUnit(40, 40, kernel_size=5, stride=1, exp_factor=6, se_factor=4,
                                                   normalization=B.lambda_batchnorm2d(), activation=B.lambda_swish(),
                                                   tf_mode=False), (2, 40, 14, 14)),
    "effi_invres_k3_s2": (lambda: M.EffiInvResUnit(24, 40, kernel_size=3, stride=2, exp_factor=6, se_factor=4,
                                                   normalization=B.lambda_batchnorm2d(), activation=B.lambda_swish(),
                                                   tf_mode=False), (1, 24, 15, 15)),
    "mnv3_unit_k5_se_hswish": (lambda: M.MobileNetV3Unit(40, 40, exp_channels=120, stride=1, use_kernel3=False,
                                                         activation=B.lambda_hswish(), use_se=True), (2, 40, 14, 14)),
    "mnv3_unit_k5_s2_se": (lambda: M.MobileNetV3Unit(24, 40, exp_channels=96, stride=2, use_kernel3=False,
                                                     activation=B.lambda_hswish(), use_se=True), (1, 24, 17, 15)),
    # asymmetric padding: ConvBlock with a 4-tuple padding (nn.ZeroPad2d, conv.py:245-249) and a tf_mode unit (F.pad)
    "convblock_3x3_s2_pad4": (lambda: B.ConvBlock(16, 24, kernel_size=3, stride=2, padding=(0, 1, 2, 1)), (2, 16, 14, 15)),
    "effi_invres_k5_s2_tf": (lambda: M.EffiInvResUnit(24, 40, kernel_size=5, stride=2, exp_factor=6, se_factor=4,
                                                      normalization=B.lambda_batchnorm2d(eps=1e-3),
                                                      activation=B.lambda_swish(), tf_mode=True), (1, 24, 16, 16)),
}


def _rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)  # the fixtures were generated single-threaded (thread count perturbs fp32 sums)
    yield
    torch.set_num_threads(n)


@pytest.mark.parametrize("stem,name,shape,sub", NETS, ids=[n[1] for n in NETS])
def test_oracle_matches_reference_golden_nets(stem, name, shape, sub):
    gold = np.load(os.path.join(GOLDEN, stem + ".npz"))
    net = seeded_init(P.get_model(name, pretrained=False).eval(), seed=0, randomize_bn=True)
    assert sum(p.numel() for p in net.parameters()) == int(gold["n_params"])
    y = oracle_forward(net, seeded_input(shape, seed=1234))
    ys = y if isinstance(y, (tuple, list)) else (y,)
    for i, t in enumerate(ys):
        g = torch.from_numpy(gold[f"out{i}"])
        t = t[..., ::sub, ::sub] if (t.dim() == 4 and sub > 1) else t
        assert t.shape == g.shape
        assert _rel(t, g) <= 1e-4, f"{name} out{i}: oracle deviates from the reference golden vector"
        if t.dim() == 2:
            assert torch.equal(t.argmax(1), g.argmax(1))


@pytest.mark.parametrize("stem", sorted(BLOCKS))
def test_oracle_matches_reference_golden_blocks(stem):
    ctor, shape = BLOCKS[stem]
    gold = torch.from_numpy(np.load(os.path.join(GOLDEN, "block_" + stem + ".npz"))["out0"])
    blk = seeded_init(ctor().eval(), seed=7, randomize_bn=True)
    y = oracle_forward(blk, seeded_input(shape, seed=99))
    assert y.shape == gold.shape
    assert _rel(y, gold) <= 1e-5


def test_state_dict_keys_match_reference():
    """Checkpoint compatibility: same keys and shapes, in the same order, as the reference's modules."""
    import hashlib
    with open(os.path.join(GOLDEN, "state_dict_keys.json")) as f:
        want = json.load(f)
    for name, rec in want.items():
        sd = P.get_model(name, pretrained=False).state_dict()
        digest = hashlib.sha1("\n".join(f"{k}:{tuple(v.shape)}" for k, v in sd.items()).encode()).hexdigest()
        assert len(sd) == rec["n"], name
        assert digest == rec["sha1"], f"{name}: state_dict keys/shapes differ from the reference"


# ---- live comparisons against the reference package (build container only) ---------------------------------------
@pytest.mark.reference
@pytest.mark.parametrize("name,shape", [("resnet18", (8, 3, 224, 224)), ("mobilenetv2_w1", (2, 3, 224, 224)),
                                        ("seresnext50_32x4d", (1, 3, 224, 224)), ("efficientnet_b0", (2, 3, 224, 224)),
                                        ("efficientnet_b1", (1, 3, 240, 240)), ("mobilenetv3_large_w1", (2, 3, 224, 224)),
                                        ("mobilenetv3_small_wd2", (1, 3, 224, 224))])
def test_oracle_equals_reference_live(reference_pkg, name, shape):
    from pytorchcv.model_provider import get_model as ref_get_model
    ref = seeded_init(ref_get_model(name, pretrained=False).eval(), seed=3, randomize_bn=True)
    x = seeded_input(shape, seed=5)
    with torch.no_grad():
        want = ref(x)
    got = oracle_forward(ref, x)           # the oracle walks the REFERENCE's own module tree ...
    assert _rel(got, want) <= 1e-6
    mine = seeded_init(P.get_model(name, pretrained=False).eval(), seed=3, randomize_bn=True)
    got2 = oracle_forward(mine, x)         # ... and the mirror tree with the same name-keyed weights
    assert _rel(got2, want) <= 1e-6


@pytest.mark.reference
@pytest.mark.parametrize("name", ["resnet18", "resnet50", "mobilenetv2_w1", "seresnext50_32x4d",
                                  "deeplabv3_resnetd50b_voc", "mobilenet_w1", "efficientnet_b0", "efficientnet_b3",
                                  "mobilenetv3_large_w1", "mobilenetv3_small_w3d4", "seresnet18", "seresnetbc26b",
                                  "fcn8sd_resnetd50b_voc", "pspnet_resnetd50b_voc", "mnasnet_b1", "mnasnet_small", "fbnet_cb", "spnasnet", "senet16", "proxylessnas_gpu", "proxylessnas_cpu"])
def test_same_seed_same_random_init_as_reference(reference_pkg, name):
    """torch.manual_seed(0); get_model(name) consumes the RNG in the reference's order -> bit-identical weights."""
    from pytorchcv.model_provider import get_model as ref_get_model
    torch.manual_seed(0)
    a = ref_get_model(name, pretrained=False).state_dict()
    torch.manual_seed(0)
    b = P.get_model(name, pretrained=False).state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k


@pytest.mark.reference
def test_reference_modules_lower_without_mirror(reference_pkg):
    """accelerate() pattern-matches on class names, so the real reference tree compiles too (dry run, no GPU)."""
    from pytorchcv.model_provider import get_model as ref_get_model
    from pytorchcv_b200 import plan as PL
    from pytorchcv_b200._lib import BF16
    net = ref_get_model("resnet50", pretrained=False).eval()
    b = PL.Builder(BF16, torch.device("cpu"))
    x = b.new(2, 224, 224, 8)
    out = PL.lower(b, net, x)
    assert (out.N, out.C, out.flat) == (2, 1000, True)
    # 53 convs + maxpool + global pool + fc, minus the projection shortcuts of stages 1-3 folded into their units' conv3
    # (pcv_conv1x1_dual; stage 4 at batch 2 is a single 128-row tile, outside the CTA-pair kernel's domain)
    assert len(b.ops) == 53


def _fold_1x1(cb):
    """A linear 1x1 ConvBlock (conv + eval-mode BatchNorm, conv.py:278-286) as (W', b') in float64."""
    c, bn = cb.conv, cb.bn
    w = c.weight.detach().double().view(c.out_channels, c.in_channels)
    b = c.bias.detach().double() if c.bias is not None else torch.zeros(c.out_channels, dtype=torch.float64)
    sc = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
    return w * sc[:, None], (b - bn.running_mean.detach().double()) * sc + bn.bias.detach().double()


@pytest.mark.parametrize("kind,cin,cout,stride", [("res", 64, 256, 1), ("res", 256, 512, 2), ("resnext", 72, 264, 2),
                                                  ("seresnext", 64, 256, 1), ("seresnext", 136, 264, 2)])
def test_projection_shortcut_is_the_second_half_of_conv3s_k_dimension(kind, cin, cout, stride):
    """The identity the dual-source conv (pcv_conv1x1_dual / _se) computes, checked against the oracle in float64:
    act(conv3(y2) + identity_conv(x)) == act([W3 | Wid] . [y2 ; x[::s]] + b3 + bid)  (ResUnit, resnet.py:221-229) and, with the
    SE gate g squeezed from conv3's output, act((W3 y2 + b3) g + Wid x[::s] + bid)  (SEResNeXtUnit, seresnext.py:57-66)."""
    if kind == "res":
        unit = M.ResUnit(cin, cout, stride=stride, bottleneck=True, conv1_stride=False)
    elif kind == "resnext":
        unit = M.ResNeXtUnit(cin, cout, stride=stride, cardinality=32, bottleneck_width=4)
    else:
        unit = M.SEResNeXtUnit(cin, cout, stride=stride, cardinality=32, bottleneck_width=4)
    unit = seeded_init(unit.eval(), seed=31, randomize_bn=True).double()
    x = seeded_input((3, cin, 13, 11), seed=32).double()
    with torch.no_grad():
        want = oracle_forward(unit, x)
        y2 = oracle_forward(unit.body.conv2, oracle_forward(unit.body.conv1, x))
        (w3, b3), (wid, bid) = _fold_1x1(unit.body.conv3), _fold_1x1(unit.identity_conv)
        xs = x[:, :, ::stride, ::stride]
        assert xs.shape[2:] == y2.shape[2:]
        if kind == "seresnext":
            main = torch.einsum("oc,nchw->nohw", w3, y2) + b3[None, :, None, None]
            # SEBlock.forward returns x * w (att.py:99-105): recover w, one value per (image, channel)
            gate = (oracle_forward(unit.se, main) * main).sum((2, 3)) / (main * main).sum((2, 3))
            got = main * gate[:, :, None, None] + torch.einsum("oc,nchw->nohw", wid, xs) + bid[None, :, None, None]
        else:
            a = torch.cat([y2, xs], dim=1)                       # [y2 ; x[::s]] along K
            w = torch.cat([w3, wid], dim=1)                      # [W3 | Wid]
            got = torch.einsum("ok,nkhw->nohw", w, a) + (b3 + bid)[None, :, None, None]
        got = torch.relu(got)
    assert got.shape == want.shape
    assert float((got - want).abs().max() / want.abs().max()) <= 1e-10
