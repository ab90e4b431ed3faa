"""ORACLE — test infrastructure only.  A CPU restatement of the reference's eval-mode forward for the hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this
module, and only as the checker or the CPU baseline; nothing under `pytorchcv_b200/` imports it.

What it restates.  osmr/pytorchcv is pure Python: each block's `forward` composes torch leaf ops, and the arithmetic
itself lives in the third-party dependency `torch` (unpinned in the reference: setup.py:32 `install_requires=
['numpy','requests','torch','torchvision']`; effective here: torch 2.11.0+cu128 CPU kernels, oneDNN conv, native
BN/ReLU/pool).  The functions below re-express every composite forward on the path — unfused, op by op, in the
reference's order (conv -> BatchNorm -> activation; body -> add -> ReLU) — over a module tree that follows the
reference's attribute names, using torch's *functional* CPU ops in fp32.  Each function cites the reference lines it
follows.

Pinning.  The reference's own tests hold no golden vectors for this path (parameter counts and output
shapes only).  The oracle is therefore pinned against outputs of the REFERENCE ITSELF run in the build container:
`tests/golden/make_golden.py` imports /root/reference, evaluates blocks and whole networks on seeded weights and
inputs, and commits the results under tests/golden/; `tests/test_oracle.py` replays them through this file
(bit-exact on the same torch build), and additionally compares oracle vs reference live whenever /root/reference is
importable.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = ["oracle_forward", "OracleUnsupported"]


class OracleUnsupported(NotImplementedError):
    pass


# ---- leaves ---------------------------------------------------------------------------------------------------------
def _activation(m: nn.Module, x: torch.Tensor) -> torch.Tensor:
    """create_activation_layer's menu (common/activ.py:188-222) as formulas."""
    name = type(m).__name__
    if isinstance(m, nn.ReLU6):
        return x.clamp(0.0, 6.0)                                   # activ.py:67-81
    if isinstance(m, nn.ReLU):
        return x.clamp_min(0.0)                                    # activ.py:50-64
    if isinstance(m, nn.Sigmoid):
        return torch.sigmoid(x)                                    # activ.py:123-132
    if name == "Swish":
        return x * torch.sigmoid(x)                                # activ.py:20-21
    if name == "HSwish":
        return x * (x + 3.0).clamp(0.0, 6.0) / 6.0                 # activ.py:46-47
    if name == "HSigmoid":
        return (x + 3.0).clamp(0.0, 6.0) / 6.0                     # activ.py:29-30
    if name == "GhostHSigmoid":
        return x.clamp(0.0, 1.0)                                   # ghostnet.py:23-24
    if isinstance(m, nn.PReLU):
        return F.prelu(x, m.weight)                                # activ.py:84-98
    if isinstance(m, nn.LeakyReLU):
        return F.leaky_relu(x, m.negative_slope)                   # activ.py:101-120
    if isinstance(m, nn.Identity):
        return x
    raise OracleUnsupported(f"activation {name}")


def _conv2d(m: nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
    """nn.Conv2d as created at conv.py:115,156,193,250."""
    return F.conv2d(x, m.weight, m.bias, stride=m.stride, padding=m.padding, dilation=m.dilation, groups=m.groups)


def _batchnorm(m: nn.BatchNorm2d, x: torch.Tensor) -> torch.Tensor:
    """Eval-mode BatchNorm2d (norm.py:34-50): (x - mean) / sqrt(var + eps) * gamma + beta with running stats."""
    return F.batch_norm(x, m.running_mean, m.running_var, m.weight, m.bias, False, 0.0, m.eps)


# ---- common blocks ---------------------------------------------------------------------------------------------------
def conv_block(m, x):
    """ConvBlock.forward (common/conv.py:278-286)."""
    if getattr(m, "use_pad", False):
        x = F.pad(x, m.pad.padding)
    x = _conv2d(m.conv, x)
    if m.normalize:
        x = _batchnorm(m.bn, x)
    if m.activate:
        x = _activation(m.activ, x)
    return x


def dws_conv_block(m, x):
    """DwsConvBlock.forward (common/conv.py:605-608)."""
    return conv_block(m.pw_conv, conv_block(m.dw_conv, x))


def se_block(m, x):
    """SEBlock.forward (common/att.py:94-105)."""
    w = F.adaptive_avg_pool2d(x, 1)
    if not m.use_conv:
        w = w.view(x.size(0), -1)
    w = _conv2d(m.conv1, w) if m.use_conv else F.linear(w, m.fc1.weight, m.fc1.bias)
    w = _activation(m.activ, w)
    w = _conv2d(m.conv2, w) if m.use_conv else F.linear(w, m.fc2.weight, m.fc2.bias)
    w = _activation(m.sigmoid, w)
    if not m.use_conv:
        w = w.unsqueeze(2).unsqueeze(3)
    return x * w


# ---- ResNet family ---------------------------------------------------------------------------------------------------
def res_body(m, x):
    """ResBlock.forward (resnet.py:63-66), ResBottleneck.forward (resnet.py:136-140), ResNeXtBottleneck.forward
    (resnext.py:56-59)."""
    x = conv_block(m.conv1, x)
    x = conv_block(m.conv2, x)
    if hasattr(m, "conv3"):
        x = conv_block(m.conv3, x)
    return x


def res_unit(m, x):
    """ResUnit.forward (resnet.py:221-229) and ResNeXtUnit.forward (resnext.py:108-116)."""
    identity = conv_block(m.identity_conv, x) if m.resize_identity else x
    x = res_body(m.body, x)
    x = x + identity
    return _activation(m.activ, x)


def se_resnext_unit(m, x):
    """SEResNeXtUnit.forward (seresnext.py:57-66) == SEResUnit.forward (seresnet.py:63-72)."""
    identity = conv_block(m.identity_conv, x) if m.resize_identity else x
    x = res_body(m.body, x)
    x = se_block(m.se, x)
    x = x + identity
    return _activation(m.activ, x)


def res_init_block(m, x):
    """ResInitBlock.forward (resnet.py:260-263): 7x7/2 ConvBlock then MaxPool2d(3, 2, 1)."""
    return _leaf(m.pool, conv_block(m.conv, x))


def se_init_block(m, x):
    """SEInitBlock.forward (senet.py:159-164)."""
    for c in (m.conv1, m.conv2, m.conv3):
        x = conv_block(c, x)
    return _leaf(m.pool, x)


def linear_bottleneck(m, x):
    """LinearBottleneck.forward (mobilenetv2.py:62-71)."""
    identity = x
    if m.use_exp_conv:
        x = conv_block(m.conv1, x)
    x = conv_block(m.conv2, x)
    x = conv_block(m.conv3, x)
    return x + identity if m.residual else x


def classifier(m, x):
    """ResNet.forward (resnet.py:333-337) == SEResNeXt.forward (seresnext.py:136-140) == ResNeXt / MobileNet."""
    x = oracle_forward(m.features, x)
    x = x.view(x.size(0), -1)
    return F.linear(x, m.output.weight, m.output.bias)


def mobilenetv2(m, x):
    """MobileNetV2.forward (mobilenetv2.py:152-156): conv1x1 classifier on the pooled 1x1 map, then view."""
    x = oracle_forward(m.features, x)
    x = _conv2d(m.output, x)
    return x.view(x.size(0), -1)


def resnetd(m, x):
    """ResNetD.forward (resnetd.py:98-106)."""
    outs = oracle_forward(m.features, x)
    if not isinstance(outs, list):
        outs = [outs]
    y = outs[0].view(outs[0].size(0), -1)
    y = F.linear(y, m.output.weight, m.output.bias)
    return [y] + outs[1:] if m.multi_output else y


# ---- containers ------------------------------------------------------------------------------------------------------
def sequential(m, x):
    for child in m.children():
        x = oracle_forward(child, x)
    return x


def concurrent(m, x):
    """Concurrent.forward (common/arch.py:84-95)."""
    outs = [oracle_forward(child, x) for child in m.children()]
    if m.merge_type == "cat":
        return torch.cat(tuple(outs), dim=m.axis)
    if m.merge_type == "stack":
        return torch.stack(tuple(outs), dim=m.axis)
    if m.merge_type == "sum":
        return torch.stack(tuple(outs), dim=m.axis).sum(m.axis)
    raise OracleUnsupported(m.merge_type)


def multi_output_sequential(m, x):
    """MultiOutputSequential.forward (common/arch.py:332-347)."""
    outs = []
    for child in m.children():
        x = oracle_forward(child, x)
        if getattr(child, "do_output", False):
            outs.append(x)
        elif getattr(child, "do_output2", False):
            outs.extend(x[1])
            x = x[0]
    if m.multi_output:
        return [x] + outs if m.return_last else outs
    if m.dual_output:
        return x, outs
    return x


# ---- DeepLabv3 -------------------------------------------------------------------------------------------------------
def deeplab_final_block(m, x, out_size):
    """DeepLabv3FinalBlock.forward (deeplabv3.py:49-54); Dropout is the identity in eval mode."""
    x = conv_block(m.conv1, x)
    x = _conv2d(m.conv2, x)
    return F.interpolate(x, size=out_size, mode="bilinear", align_corners=True)


def aspp_avg_branch(m, x):
    """ASPPAvgBranch.forward (deeplabv3.py:80-87)."""
    in_size = m.upscale_out_size if m.upscale_out_size is not None else x.shape[2:]
    x = F.adaptive_avg_pool2d(x, 1)
    x = conv_block(m.conv, x)
    return F.interpolate(x, size=in_size, mode="bilinear", align_corners=True)


def aspp(m, x):
    """AtrousSpatialPyramidPooling.forward (deeplabv3.py:129-133)."""
    return conv_block(m.conv, concurrent(m.branches, x))


def deeplabv3(m, x):
    """DeepLabv3.forward (deeplabv3.py:199-208)."""
    in_size = m.in_size if m.fixed_size else x.shape[2:]
    x, y = multi_output_sequential(m.backbone, x)
    x = aspp(m.pool, x)
    x = deeplab_final_block(m.final_block, x, in_size)
    if m.aux:
        return x, deeplab_final_block(m.aux_block, y, in_size)
    return x


# ---- EfficientNet -----------------------------------------------------------------------------------
def _tf_pad(x, kernel_size, stride=1, dilation=1):
    """calc_tf_padding (efficientnet.py:27-55) applied with F.pad, as the tf_mode forwards do."""
    import math
    h, w = x.shape[2:]
    oh, ow = math.ceil(h / stride), math.ceil(w / stride)
    ph = max((oh - 1) * stride + (kernel_size - 1) * dilation + 1 - h, 0)
    pw = max((ow - 1) * stride + (kernel_size - 1) * dilation + 1 - w, 0)
    return F.pad(x, (ph // 2, ph - ph // 2, pw // 2, pw - pw // 2))


def effi_init_block(m, x):
    """EffiInitBlock.forward (efficientnet.py:235-239)."""
    if m.tf_mode:
        x = _tf_pad(x, 3, 2)
    return conv_block(m.conv, x)


def effi_dws_conv_unit(m, x):
    """EffiDwsConvUnit.forward (efficientnet.py:105-115)."""
    identity = x
    if m.tf_mode:
        x = _tf_pad(x, 3)
    x = conv_block(m.pw_conv, se_block(m.se, conv_block(m.dw_conv, x)))
    return x + identity if m.residual else x


def effi_inv_res_unit(m, x):
    """EffiInvResUnit.forward (efficientnet.py:185-197)."""
    identity = x
    x = conv_block(m.conv1, x)
    if m.tf_mode:
        x = _tf_pad(x, m.kernel_size, m.stride)
    x = conv_block(m.conv2, x)
    if m.use_se:
        x = se_block(m.se, x)
    x = conv_block(m.conv3, x)
    return x + identity if m.residual else x


def mobilenetv3_unit(m, x):
    """MobileNetV3Unit.forward (mobilenetv3.py:82-93)."""
    identity = x
    if m.use_exp_conv:
        x = conv_block(m.exp_conv, x)
    x = conv_block(m.conv1, x)
    if m.use_se:
        x = se_block(m.se, x)
    x = conv_block(m.conv2, x)
    return x + identity if m.residual else x


def dws_exp_se_res_unit(m, x):
    """DwsExpSEResUnit.forward (mnasnet.py:77-88)."""
    identity = x
    if m.use_exp_conv:
        x = conv_block(m.exp_conv, x)
    x = conv_block(m.dw_conv, x)
    if m.use_se:
        x = se_block(m.se, x)
    x = conv_block(m.pw_conv, x)
    return x + identity if m.residual else x


def fbnet_unit(m, x):
    """FBNetUnit.forward (fbnet.py:77-87) == SPNASUnit.forward (spnasnet.py:72-82)."""
    identity = x
    if m.use_exp_conv:
        x = conv_block(m.exp_conv, x)
    x = conv_block(m.conv2, conv_block(m.conv1, x))
    return x + identity if m.residual else x


def proxyless_unit(m, x):
    """ProxylessUnit.forward (proxylessnas.py:114-123) with ProxylessBlock.forward (proxylessnas.py:64-70)."""
    if not m.residual:
        return x
    y = conv_block(m.body.bc_conv, x) if m.body.use_bc else x
    y = conv_block(m.body.pw_conv, conv_block(m.body.dw_conv, y))
    return x + y if m.shortcut else y


def mnas_edge_block(m, x):
    """MnasInitBlock.forward / MnasFinalBlock.forward (mnasnet.py:121-124, 157-160)."""
    return oracle_forward(m.conv2, oracle_forward(m.conv1, x))


def mobilenetv3_final_block(m, x):
    """MobileNetV3FinalBlock.forward (mobilenetv3.py:127-131)."""
    x = conv_block(m.conv, x)
    return se_block(m.se, x) if m.use_se else x


def mobilenetv3_classifier(m, x):
    """MobileNetV3Classifier.forward (mobilenetv3.py:168-174); Dropout is the identity in eval."""
    return _conv2d(m.conv2, _activation(m.activ, _conv2d(m.conv1, x)))


def mobilenetv3(m, x):
    """MobileNetV3.forward (mobilenetv3.py:277-281)."""
    x = mobilenetv3_classifier(m.output, oracle_forward(m.features, x))
    return x.view(x.size(0), -1)


def efficientnet(m, x):
    """EfficientNet.forward (efficientnet.py:354-358): features -> view -> output (Dropout is the identity in eval)."""
    x = oracle_forward(m.features, x)
    x = x.view(x.size(0), -1)
    return sequential(m.output, x)


def pyramid_pooling_branch(m, x):
    """PyramidPoolingBranch.forward (pspnet.py:71-75)."""
    in_size = m.upscale_out_size if m.upscale_out_size is not None else x.shape[2:]
    x = conv_block(m.conv, F.adaptive_avg_pool2d(x, m.pool.output_size))
    return F.interpolate(x, size=in_size, mode="bilinear", align_corners=True)


def pyramid_pooling(m, x):
    """PyramidPooling.forward (pspnet.py:116-118): Concurrent(Identity, 4 pooled branches), concatenated on channels."""
    return concurrent(m.branches, x)


def fcn8sd(m, x):
    """FCN8sd.forward (fcn8sd.py:112-120); FCNFinalBlock.forward (fcn8sd.py:47-52) has DeepLabv3FinalBlock's body.
    PSPNet.forward (pspnet.py:196-205) is the same with the pyramid pooling module before the head."""
    in_size = m.in_size if m.fixed_size else x.shape[2:]
    x, y = multi_output_sequential(m.backbone, x)
    if hasattr(m, "pool"):
        x = pyramid_pooling(m.pool, x)
    x = deeplab_final_block(m.final_block, x, in_size)
    if m.aux:
        return x, deeplab_final_block(m.aux_block, y, in_size)
    return x


# ---- pre-activation family, DarkNet-53 -------------------------------------------------------
def pre_conv_block(m, x):
    """PreConvBlock.forward (common/conv.py:717-731): BN -> activation -> conv; optionally also the pre-activated map."""
    if m.normalize:
        x = _batchnorm(m.bn, x)
    if m.activate:
        x = _activation(m.activ, x)
    y = _conv2d(m.conv, x)
    return (y, x) if m.return_preact else y


def pre_res_body(m, x):
    """PreResBlock.forward (preresnet.py:56-59) / PreResBottleneck.forward (preresnet.py:98-102)."""
    x, x_pre_activ = pre_conv_block(m.conv1, x)
    x = pre_conv_block(m.conv2, x)
    if hasattr(m, "conv3"):
        x = pre_conv_block(m.conv3, x)
    return x, x_pre_activ


def pre_res_unit(m, x):
    """PreResUnit.forward (preresnet.py:157-163)."""
    identity = x
    x, x_pre_activ = pre_res_body(m.body, x)
    if m.resize_identity:
        identity = _conv2d(m.identity_conv, x_pre_activ)
    return x + identity


def pre_res_init_block(m, x):
    """PreResInitBlock.forward (preresnet.py:195-200)."""
    return _leaf(m.pool, _activation(m.activ, _batchnorm(m.bn, _conv2d(m.conv, x))))


def pre_res_activation(m, x):
    """PreResActivation.forward (preresnet.py:218-221)."""
    return _activation(m.activ, _batchnorm(m.bn, x))


def ghost_conv_block(m, x):
    """GhostConvBlock.forward (ghostnet.py:57-60)."""
    x = conv_block(m.main_conv, x)
    y = conv_block(m.cheap_conv, x)
    return torch.cat((x, y), dim=1)


def ghost_exp_block(m, x):
    """GhostExpBlock.forward (ghostnet.py:114-121)."""
    x = ghost_conv_block(m.exp_conv, x)
    if m.use_dw_conv:
        x = conv_block(m.dw_conv, x)
    if m.use_se:
        x = se_block(m.se, x)
    return ghost_conv_block(m.pw_conv, x)


def ghost_unit(m, x):
    """GhostUnit.forward (ghostnet.py:167-174)."""
    identity = dws_conv_block(m.identity_conv, x) if m.resize_identity else x
    return ghost_exp_block(m.body, x) + identity


def ghostnet(m, x):
    """GhostNet.forward (ghostnet.py:298-302) with GhostClassifier.forward (ghostnet.py:203-206)."""
    x = oracle_forward(m.features, x)
    x = _conv2d(m.output.conv2, conv_block(m.output.conv1, x))
    return x.view(x.size(0), -1)


def mix_conv_block(m, x):
    """MixConvBlock.forward (mixnet.py:151-157) with MixConv.forward (mixnet.py:76-80)."""
    xx = torch.split(x, m.conv.splitted_in_channels, dim=m.conv.axis)
    x = torch.cat(tuple(_conv2d(conv_i, x_i) for x_i, conv_i in zip(xx, m.conv.children())), dim=m.conv.axis)
    if m.normalize:
        x = _batchnorm(m.bn, x)
    if m.activate:
        x = _activation(m.activ, x)
    return x


def mix_unit(m, x):
    """MixUnit.forward (mixnet.py:282-293)."""
    identity = x
    if m.use_exp_conv:
        x = oracle_forward(m.exp_conv, x)
    x = oracle_forward(m.conv1, x)
    if m.use_se:
        x = se_block(m.se, x)
    x = oracle_forward(m.conv2, x)
    return x + identity if m.residual else x


def mix_init_block(m, x):
    """MixInitBlock.forward (mixnet.py:326-329)."""
    return mix_unit(m.conv2, conv_block(m.conv1, x))


def effi_edge_res_unit(m, x):
    """EffiEdgeResUnit.forward (efficientnetedge.py:77-86)."""
    identity = x
    x = conv_block(m.conv1, x)
    if m.use_se:
        x = se_block(m.se, x)
    x = conv_block(m.conv2, x)
    return x + identity if m.residual else x


def dark_unit(m, x):
    """DarkUnit.forward (darknet53.py:45-49)."""
    return conv_block(m.conv2, conv_block(m.conv1, x)) + x


# ---- dispatch --------------------------------------------------------------------------------------------------------
def _leaf(m, x):
    if isinstance(m, nn.Conv2d):
        return _conv2d(m, x)
    if isinstance(m, nn.BatchNorm2d):
        return _batchnorm(m, x)
    if isinstance(m, nn.MaxPool2d):                                  # resnet.py:255-258, senet.py:154-157
        return F.max_pool2d(x, m.kernel_size, m.stride, m.padding, m.dilation, m.ceil_mode)
    if isinstance(m, nn.AvgPool2d):                                  # resnet.py:316-318
        return F.avg_pool2d(x, m.kernel_size, m.stride, m.padding, m.ceil_mode, m.count_include_pad)
    if isinstance(m, nn.AdaptiveAvgPool2d):                          # att.py:72, resnetd.py:83
        return F.adaptive_avg_pool2d(x, m.output_size)
    if isinstance(m, nn.Linear):
        return F.linear(x, m.weight, m.bias)
    if isinstance(m, (nn.Dropout, nn.Identity)):
        return x
    return _activation(m, x)


_BY_NAME = {
    "ConvBlock": conv_block, "DwsConvBlock": dws_conv_block, "SEBlock": se_block,
    "ResBlock": res_body, "ResBottleneck": res_body, "ResNeXtBottleneck": res_body, "SENetBottleneck": res_body,
    "SENetUnit": se_resnext_unit, "SENet": efficientnet,
    "ResUnit": res_unit, "ResNeXtUnit": res_unit, "SEResNeXtUnit": se_resnext_unit, "SEResUnit": se_resnext_unit,
    "SEResNet": classifier, "FCN8sd": fcn8sd, "PSPNet": fcn8sd, "PyramidPoolingBranch": pyramid_pooling_branch,
    "PyramidPooling": pyramid_pooling, "Identity": lambda m, x: x,
    "ResInitBlock": res_init_block, "SEInitBlock": se_init_block, "LinearBottleneck": linear_bottleneck,
    "ResNet": classifier, "SEResNeXt": classifier, "ResNeXt": classifier, "MobileNet": classifier,
    "MobileNetV2": mobilenetv2, "ResNetD": resnetd,
    "EffiInitBlock": effi_init_block, "EffiDwsConvUnit": effi_dws_conv_unit, "EffiInvResUnit": effi_inv_res_unit,
    "EfficientNet": efficientnet,
    "DwsExpSEResUnit": dws_exp_se_res_unit, "MnasInitBlock": mnas_edge_block, "MnasFinalBlock": mnas_edge_block,
    "MnasNet": classifier, "FBNetUnit": fbnet_unit, "FBNetInitBlock": mnas_edge_block, "FBNet": classifier, "ProxylessUnit": proxyless_unit, "ProxylessNAS": classifier,
    "SPNASUnit": fbnet_unit, "SPNASInitBlock": mnas_edge_block, "SPNASFinalBlock": mnas_edge_block, "SPNASNet": classifier,
    "MobileNetV3Unit": mobilenetv3_unit, "MobileNetV3FinalBlock": mobilenetv3_final_block,
    "MobileNetV3Classifier": mobilenetv3_classifier, "MobileNetV3": mobilenetv3,
    "PreConvBlock": pre_conv_block, "PreResBlock": pre_res_body, "PreResBottleneck": pre_res_body, "PreResUnit": pre_res_unit,
    "PreResInitBlock": pre_res_init_block, "PreResActivation": pre_res_activation, "PreResNet": classifier,
    "DarkUnit": dark_unit, "DarkNet53": classifier,
    "EffiEdgeResUnit": effi_edge_res_unit, "EfficientNetEdge": efficientnet,
    "MixConvBlock": mix_conv_block, "MixUnit": mix_unit, "MixInitBlock": mix_init_block, "MixNet": classifier,
    "GhostConvBlock": ghost_conv_block, "GhostExpBlock": ghost_exp_block, "GhostUnit": ghost_unit, "GhostNet": ghostnet,
    "Concurrent": concurrent, "MultiOutputSequential": multi_output_sequential,
    "ASPPAvgBranch": aspp_avg_branch, "AtrousSpatialPyramidPooling": aspp, "DeepLabv3": deeplabv3,
}


@torch.no_grad()
def oracle_forward(m: nn.Module, x: torch.Tensor, **kw):
    """Evaluate module tree `m` on CPU fp32 input `x` the way the reference's forward would."""
    if x.is_cuda:
        raise ValueError("the oracle is a CPU restatement; pass CPU tensors")
    name = type(m).__name__
    if name in ("DeepLabv3FinalBlock", "FCNFinalBlock", "PSPFinalBlock"):
        return deeplab_final_block(m, x, kw["out_size"])
    fn = _BY_NAME.get(name)
    if fn is not None:
        return fn(m, x)
    if type(m) is nn.Sequential:
        return sequential(m, x)
    try:
        return _leaf(m, x)
    except OracleUnsupported:
        raise OracleUnsupported(f"module {name} is not on the hot path the oracle restates") from None
