"""Generates tests/golden/*.npz by running the REFERENCE ITSELF (osmr/pytorchcv at /root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):
    PYTHONPATH=/root/reference:. python tests/golden/make_golden.py
The vectors pin `oracle/` (tests/test_oracle.py) and, through it, the CUDA path (tests/test_gpu_*.py).
Weights/inputs are not stored: they are regenerated from seeds by oracle/seeded.py (name-keyed, order-independent).
"""
import hashlib
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from pytorchcv.model_provider import get_model as ref_get_model  # noqa: E402
from pytorchcv.models.common.conv import ConvBlock, DwsConvBlock, conv3x3_block, dwconv5x5_block  # noqa: E402
from pytorchcv.models.common.att import SEBlock  # noqa: E402
from pytorchcv.models.resnet import ResUnit  # noqa: E402
from pytorchcv.models.mobilenetv2 import LinearBottleneck  # noqa: E402
from pytorchcv.models.seresnext import SEResNeXtUnit  # noqa: E402
from pytorchcv.models.common.activ import lambda_relu6, lambda_swish, lambda_hswish, lambda_prelu, lambda_leakyrelu  # noqa: E402
from pytorchcv.models.common.conv import PreConvBlock, dwconv3x3_block  # noqa: E402
from pytorchcv.models.preresnet import PreResUnit  # noqa: E402
from pytorchcv.models.ghostnet import GhostConvBlock, GhostUnit  # noqa: E402
from pytorchcv.models.mixnet import MixConvBlock, MixUnit, mixconv1x1_block  # noqa: E402
from pytorchcv.models.mobilenetv3 import MobileNetV3Unit  # noqa: E402
from pytorchcv.models.common.norm import lambda_batchnorm2d  # noqa: E402
from pytorchcv.models.efficientnet import EffiDwsConvUnit, EffiInvResUnit  # noqa: E402

from oracle.seeded import seeded_init, seeded_input  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# (file stem, model name, kwargs, input shape, seed, output subsample stride for dense maps)
NETS = [
    ("resnet18_bs2", "resnet18", {}, (2, 3, 224, 224), 0, 1),
    ("resnet50_bs2", "resnet50", {}, (2, 3, 224, 224), 0, 1),
    ("mobilenetv2_w1_bs2", "mobilenetv2_w1", {}, (2, 3, 224, 224), 0, 1),
    ("seresnext50_32x4d_bs2", "seresnext50_32x4d", {}, (2, 3, 224, 224), 0, 1),
    ("mobilenet_w1_bs2", "mobilenet_w1", {}, (2, 3, 224, 224), 0, 1),
    ("deeplabv3_resnetd50b_voc_bs1", "deeplabv3_resnetd50b_voc", {}, (1, 3, 480, 480), 0, 16),
    ("efficientnet_b0_bs2", "efficientnet_b0", {}, (2, 3, 224, 224), 0, 1),
    ("mobilenetv3_large_w1_bs2", "mobilenetv3_large_w1", {}, (2, 3, 224, 224), 0, 1),
    ("mobilenetv3_small_w1_bs2", "mobilenetv3_small_w1", {}, (2, 3, 224, 224), 0, 1),
    ("seresnet18_bs2", "seresnet18", {}, (2, 3, 224, 224), 0, 1),
    ("seresnet50_bs2", "seresnet50", {}, (2, 3, 224, 224), 0, 1),
    ("fcn8sd_resnetd50b_voc_bs1", "fcn8sd_resnetd50b_voc", {}, (1, 3, 480, 480), 0, 16),
    ("pspnet_resnetd50b_voc_bs1", "pspnet_resnetd50b_voc", {}, (1, 3, 480, 480), 0, 16),
    ("mnasnet_a1_bs2", "mnasnet_a1", {}, (2, 3, 224, 224), 0, 1),
    ("mnasnet_small_bs2", "mnasnet_small", {}, (2, 3, 224, 224), 0, 1),
    ("fbnet_cb_bs2", "fbnet_cb", {}, (2, 3, 224, 224), 0, 1),
    ("spnasnet_bs2", "spnasnet", {}, (2, 3, 224, 224), 0, 1),
    ("senet16_bs2", "senet16", {}, (2, 3, 224, 224), 0, 1),
    ("proxylessnas_mobile_bs2", "proxylessnas_mobile", {}, (2, 3, 224, 224), 0, 1),
    ("efficientnet_b0b_bs2", "efficientnet_b0b", {}, (2, 3, 224, 224), 0, 1),    # tf_mode: asymmetric "SAME" padding
    ("preresnet18_bs2", "preresnet18", {}, (2, 3, 224, 224), 0, 1),              # PreConvBlock family
    ("preresnet50_bs2", "preresnet50", {}, (2, 3, 224, 224), 0, 1),
    ("darknet53_bs2", "darknet53", {}, (2, 3, 224, 224), 0, 1),                  # LeakyReLU epilogues
    ("ghostnet_bs2", "ghostnet", {}, (2, 3, 224, 224), 0, 1),                    # torch.cat of odd-width halves
    ("mixnet_s_bs2", "mixnet_s", {}, (2, 3, 224, 224), 0, 1),                    # torch.split / mixed depthwise kernels 3..11
    ("efficientnet_edge_small_b_bs2", "efficientnet_edge_small_b", {}, (2, 3, 224, 224), 0, 1),   # EffiEdgeResUnit, tf_mode
]

NO_MIRROR = {"preresnet18", "preresnet50", "darknet53", "ghostnet", "mixnet_s", "efficientnet_edge_small_b"}

# block-level cases: (stem, ctor, input shape)
BLOCKS = [
    ("convblock_3x3_s2", lambda: conv3x3_block(in_channels=16, out_channels=24, stride=2), (2, 16, 15, 15)),
    ("convblock_1x1_noact", lambda: ConvBlock(32, 64, kernel_size=1, activation=None), (2, 32, 9, 9)),
    ("convblock_3x3_d2_bias", lambda: ConvBlock(16, 16, kernel_size=3, padding=2, dilation=2, bias=True), (1, 16, 12, 12)),
    ("dws_3x3", lambda: DwsConvBlock(16, 32, kernel_size=3, stride=1, padding=1), (2, 16, 10, 10)),
    ("dwconv5x5_relu6", lambda: dwconv5x5_block(in_channels=24, out_channels=24, activation=lambda_relu6()), (1, 24, 11, 11)),
    ("seblock_64", lambda: SEBlock(channels=64), (2, 64, 7, 7)),
    ("resunit_bottleneck_s2", lambda: ResUnit(64, 128, stride=2, bottleneck=True, conv1_stride=True), (2, 64, 14, 14)),
    ("resunit_basic", lambda: ResUnit(32, 32, stride=1, bottleneck=False), (2, 32, 8, 8)),
    ("linear_bottleneck_res", lambda: LinearBottleneck(24, 24, stride=1, expansion=True, remove_exp_conv=False,
                                                       activation=lambda_relu6()), (2, 24, 14, 14)),
    ("seresnext_unit", lambda: SEResNeXtUnit(256, 256, stride=1, cardinality=32, bottleneck_width=4), (1, 256, 8, 8)),
    ("effi_dws_unit", lambda: EffiDwsConvUnit(32, 16, stride=1, normalization=lambda_batchnorm2d(),
                                              activation=lambda_swish(), tf_mode=False), (2, 32, 16, 16)),
    ("effi_invres_k5_se", lambda: EffiInvResUnit(40, 40, kernel_size=5, stride=1, exp_factor=6, se_factor=4,
                                                 normalization=lambda_batchnorm2d(), activation=lambda_swish(),
                                                 tf_mode=False), (2, 40, 14, 14)),
    ("effi_invres_k3_s2", lambda: EffiInvResUnit(24, 40, kernel_size=3, stride=2, exp_factor=6, se_factor=4,
                                                 normalization=lambda_batchnorm2d(), activation=lambda_swish(),
                                                 tf_mode=False), (1, 24, 15, 15)),
    ("mnv3_unit_k5_se_hswish", lambda: MobileNetV3Unit(40, 40, exp_channels=120, stride=1, use_kernel3=False,
                                                       activation=lambda_hswish(), use_se=True), (2, 40, 14, 14)),
    ("convblock_3x3_s2_pad4", lambda: ConvBlock(16, 24, kernel_size=3, stride=2, padding=(0, 1, 2, 1)), (2, 16, 14, 15)),
    ("effi_invres_k5_s2_tf", lambda: EffiInvResUnit(24, 40, kernel_size=5, stride=2, exp_factor=6, se_factor=4,
                                                    normalization=lambda_batchnorm2d(eps=1e-3), activation=lambda_swish(),
                                                    tf_mode=True), (1, 24, 16, 16)),
    ("mnv3_unit_k5_s2_se", lambda: MobileNetV3Unit(24, 40, exp_channels=96, stride=2, use_kernel3=False,
                                                   activation=lambda_hswish(), use_se=True), (1, 24, 17, 15)),
    ("convblock_3x3_prelu", lambda: conv3x3_block(in_channels=16, out_channels=24, activation=lambda_prelu(24)), (2, 16, 13, 13)),
    ("convblock_1x1_prelu1", lambda: ConvBlock(32, 64, kernel_size=1, activation=lambda_prelu(1)), (2, 32, 9, 9)),
    ("convblock_3x3_leaky", lambda: conv3x3_block(in_channels=16, out_channels=32, stride=2,
                                                  activation=lambda_leakyrelu(negative_slope=0.1)), (2, 16, 15, 15)),
    ("dwconv3x3_leaky", lambda: dwconv3x3_block(in_channels=24, out_channels=24,
                                                activation=lambda_leakyrelu(negative_slope=0.2)), (1, 24, 11, 11)),
    ("preconv_3x3_preact", lambda: PreConvBlock(16, 32, kernel_size=3, stride=1, padding=1, return_preact=True), (2, 16, 12, 12)),
    ("preconv_1x1_s2_bias", lambda: PreConvBlock(24, 16, kernel_size=1, stride=2, padding=0, bias=True), (2, 24, 10, 10)),
    ("preresunit_bottleneck_s2", lambda: PreResUnit(64, 128, stride=2, bottleneck=True, conv1_stride=True), (2, 64, 14, 14)),
    ("preresunit_basic", lambda: PreResUnit(32, 32, stride=1, bottleneck=False, conv1_stride=False), (2, 32, 8, 8)),
    ("ghostconv_24_72", lambda: GhostConvBlock(24, 72), (2, 24, 14, 14)),                     # halves of 36: padded to 40
    ("ghostunit_16_24_s2", lambda: GhostUnit(16, 24, stride=2, use_kernel3=True, exp_factor=3.0, use_se=False), (2, 16, 28, 28)),
    ("ghostunit_24_24", lambda: GhostUnit(24, 24, stride=1, use_kernel3=True, exp_factor=3.0, use_se=False), (2, 24, 14, 14)),
    ("ghostunit_24_40_s2_k5_se", lambda: GhostUnit(24, 40, stride=2, use_kernel3=False, exp_factor=3.0, use_se=True), (2, 24, 28, 28)),
    ("ghostunit_80_80_se", lambda: GhostUnit(80, 80, stride=1, use_kernel3=True, exp_factor=2.3, use_se=True), (1, 80, 14, 14)),
    ("mixconv_dw_240_k4_s2", lambda: MixConvBlock(240, 240, kernel_size=[3, 5, 7, 9], stride=2, padding=[1, 2, 3, 4], groups=240,
                                                  activation=lambda_swish()), (1, 240, 28, 28)),      # 4 parts of 60: padded to 64
    ("mixconv1x1_40_120_k2", lambda: mixconv1x1_block(in_channels=40, out_channels=120, kernel_count=2), (2, 40, 14, 14)),
    ("mixunit_40_40_se", lambda: MixUnit(40, 40, stride=1, exp_kernel_count=2, conv1_kernel_count=2, conv2_kernel_count=2,
                                         exp_factor=6, se_factor=2, activation=lambda_swish()), (2, 40, 14, 14)),
    ("mixunit_24_40_s2_k3", lambda: MixUnit(24, 40, stride=2, exp_kernel_count=1, conv1_kernel_count=3, conv2_kernel_count=1,
                                            exp_factor=6, se_factor=2, activation=lambda_swish()), (1, 24, 28, 28)),
]


def sha(a: np.ndarray) -> str:
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


@torch.no_grad()
def main():
    torch.set_num_threads(1)  # thread count perturbs fp32 sums; fix it for reproducible fixtures
    only = sys.argv[1] if len(sys.argv) > 1 else ""   # optional substring: (re)generate only the matching fixtures
    keys = {}
    keys_path = os.path.join(OUT, "state_dict_keys.json")
    if only and os.path.exists(keys_path):
        import json
        keys = json.load(open(keys_path))
    for stem, name, kw, shape, seed, sub in NETS:
        if only not in stem:
            continue
        net = seeded_init(ref_get_model(name, pretrained=False, **kw).eval(), seed=seed, randomize_bn=True)
        x = seeded_input(shape, seed=1234)
        y = net(x)
        ys = y if isinstance(y, (tuple, list)) else (y,)
        arrs = {}
        for i, t in enumerate(ys):
            a = t.numpy()
            arrs[f"out{i}_sha1"] = np.frombuffer(sha(a).encode(), dtype=np.uint8)
            arrs[f"out{i}"] = a[..., ::sub, ::sub] if (a.ndim == 4 and sub > 1) else a
        arrs["n_params"] = np.array(sum(p.numel() for p in net.parameters()))
        if name in NO_MIRROR:   # lowered from the reference's own modules: no mirror class whose state_dict to pin
            np.savez_compressed(os.path.join(OUT, stem + ".npz"), **arrs)
            print(stem, [tuple(t.shape) for t in ys], int(arrs["n_params"]))
            continue
        keys[name] = {"n": len(net.state_dict()),
                      "sha1": hashlib.sha1("\n".join(f"{k}:{tuple(v.shape)}" for k, v in net.state_dict().items())
                                           .encode()).hexdigest()}
        np.savez_compressed(os.path.join(OUT, stem + ".npz"), **arrs)
        print(stem, [tuple(t.shape) for t in ys], int(arrs["n_params"]))
    import json
    with open(os.path.join(OUT, "state_dict_keys.json"), "w") as f:
        json.dump(keys, f, indent=1, sort_keys=True)
    for stem, ctor, shape in BLOCKS:
        if only not in stem:
            continue
        blk = seeded_init(ctor().eval(), seed=7, randomize_bn=True)
        x = seeded_input(shape, seed=99)
        y = blk(x)
        ys = y if isinstance(y, (tuple, list)) else (y,)
        np.savez_compressed(os.path.join(OUT, "block_" + stem + ".npz"), **{f"out{i}": t.numpy() for i, t in enumerate(ys)})
        print("block", stem, [tuple(t.shape) for t in ys])


if __name__ == "__main__":
    main()
