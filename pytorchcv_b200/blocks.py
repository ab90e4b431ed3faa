"""Host-side mirror of the reference's shared building blocks (pytorchcv/models/common/{conv,att,activ,norm}.py).

Same names, constructor arguments, attribute names and state_dict keys as the reference, so checkpoints load
unchanged and `get_model(name, pretrained=False)` is a drop-in.  The modules only HOLD parameters (nn.Conv2d /
nn.BatchNorm2d leaves are never called): every `forward` compiles the block into fused sm_100a kernels
(`plan.run_module`) and runs them through libpcv_b200.so.  There is no torch-op or CPU fallback.
"""
from __future__ import annotations

from inspect import isfunction

import torch.nn as nn

from .plan import run_module

__all__ = [
    "B200Module", "Swish", "HSwish", "HSigmoid", "lambda_relu", "lambda_relu6", "lambda_sigmoid", "lambda_swish",
    "lambda_hswish", "lambda_hsigmoid", "create_activation_layer", "lambda_batchnorm2d",
    "create_normalization_layer", "round_channels", "conv1x1", "conv3x3", "depthwise_conv3x3", "ConvBlock",
    "conv1x1_block", "conv3x3_block", "conv5x5_block", "conv7x7_block", "dwconv_block", "dwconv3x3_block",
    "dwconv5x5_block", "DwsConvBlock", "dwsconv3x3_block", "SEBlock",
]


class B200Module(nn.Module):
    """Base of every mirror block: forward(x) = compile-once-per-shape + run on the B200 path."""

    def forward(self, x):
        return run_module(self, x)


# ---- activations (activ.py) : parameter-free markers, fused into the producing kernel's epilogue ----------------
class _FusedActivation(nn.Module):
    def forward(self, x):
        raise RuntimeError(f"{type(self).__name__} is fused into the preceding convolution's epilogue on the B200 "
                           "path and is not callable on its own")


class Swish(_FusedActivation):
    """x * sigmoid(x) (activ.py:16-21)."""


class HSigmoid(_FusedActivation):
    """relu6(x + 3) / 6 (activ.py:24-30)."""


class HSwish(_FusedActivation):
    """x * relu6(x + 3) / 6 (activ.py:33-47)."""

    def __init__(self, inplace: bool = False):
        super().__init__()
        self.inplace = inplace


def lambda_relu(inplace: bool = True):
    return lambda: nn.ReLU(inplace=inplace)


def lambda_relu6(inplace: bool = True):
    return lambda: nn.ReLU6(inplace=inplace)


def lambda_sigmoid():
    return lambda: nn.Sigmoid()


def lambda_swish():
    return lambda: Swish()


def lambda_hswish(inplace: bool = True):
    return lambda: HSwish(inplace=inplace)


def lambda_hsigmoid():
    return lambda: HSigmoid()


_ACT_BY_NAME = {
    "relu": lambda: nn.ReLU(inplace=True),
    "relu6": lambda: nn.ReLU6(inplace=True),
    "swish": Swish,
    "hswish": lambda: HSwish(inplace=True),
    "sigmoid": nn.Sigmoid,
    "hsigmoid": HSigmoid,
}


def create_activation_layer(activation):
    """function / str / module -> activation module; unknown strings raise NotImplementedError (activ.py:188-222)."""
    assert activation is not None
    if isfunction(activation):
        return activation()
    if isinstance(activation, str):
        if activation not in _ACT_BY_NAME:
            raise NotImplementedError()
        return _ACT_BY_NAME[activation]()
    assert isinstance(activation, nn.Module)
    return activation


# ---- normalisation (norm.py) ------------------------------------------------------------------------------------
def lambda_batchnorm2d(eps: float = 1e-5):
    return lambda num_features: nn.BatchNorm2d(num_features=num_features, eps=eps)


def create_normalization_layer(normalization, **kwargs):
    """function / module -> normalisation module (norm.py:95-115)."""
    assert normalization is not None
    if isfunction(normalization):
        return normalization(**kwargs)
    assert isinstance(normalization, nn.Module)
    return normalization


def round_channels(channels, divisor: int = 8) -> int:
    """Make-divisible rounding (att.py:15-35)."""
    rounded = max(int(channels + divisor / 2.0) // divisor * divisor, divisor)
    return rounded + divisor if float(rounded) < 0.9 * channels else rounded


# ---- bare convolutions (conv.py:89-201): plain nn.Conv2d parameter holders ---------------------------------------
def conv1x1(in_channels, out_channels, stride=1, groups=1, bias=False):
    return nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=stride, groups=groups, bias=bias)


def conv3x3(in_channels, out_channels, stride=1, padding=1, dilation=1, groups=1, bias=False):
    return nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=stride, padding=padding, dilation=dilation,
                     groups=groups, bias=bias)


def depthwise_conv3x3(channels, stride=1, padding=1, dilation=1, bias=False):
    return nn.Conv2d(channels, channels, kernel_size=3, stride=stride, padding=padding, dilation=dilation,
                     groups=channels, bias=bias)


# ---- ConvBlock and factories (conv.py:204-543) -----------------------------------------------------------------
class ConvBlock(B200Module):
    """conv -> BatchNorm -> activation as ONE fused kernel; attributes .conv/.bn/.activ as in conv.py:231-276."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=False,
                 normalization=lambda_batchnorm2d(), activation=lambda_relu()):
        super().__init__()
        self.normalize = normalization is not None
        self.activate = activation is not None
        self.use_pad = isinstance(padding, (list, tuple)) and len(padding) == 4
        if self.use_pad:
            self.pad = nn.ZeroPad2d(padding=padding)
            padding = 0
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                              dilation=dilation, groups=groups, bias=bias)
        if self.normalize:
            self.bn = create_normalization_layer(normalization=normalization, num_features=out_channels)
            if self.bn is None:
                self.normalize = False
            else:
                assert isinstance(self.bn, nn.Module)
        if self.activate:
            self.activ = create_activation_layer(activation)
            if self.activ is None:
                self.activate = False
            else:
                assert isinstance(self.activ, nn.Module)


def conv1x1_block(padding=0, **kwargs):
    return ConvBlock(kernel_size=1, padding=padding, **kwargs)


def conv3x3_block(padding=1, **kwargs):
    return ConvBlock(kernel_size=3, padding=padding, **kwargs)


def conv5x5_block(padding=2, **kwargs):
    return ConvBlock(kernel_size=5, padding=padding, **kwargs)


def conv7x7_block(padding=3, **kwargs):
    return ConvBlock(kernel_size=7, padding=padding, **kwargs)


def dwconv_block(out_channels, padding=1, **kwargs):
    return ConvBlock(out_channels=out_channels, padding=padding, groups=out_channels, **kwargs)


def dwconv3x3_block(padding=1, **kwargs):
    return dwconv_block(kernel_size=3, padding=padding, **kwargs)


def dwconv5x5_block(padding=2, **kwargs):
    return dwconv_block(kernel_size=5, padding=padding, **kwargs)


class DwsConvBlock(B200Module):
    """Depthwise ConvBlock then pointwise ConvBlock (conv.py:546-608)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, bias=False,
                 dw_normalization=lambda_batchnorm2d(), pw_normalization=lambda_batchnorm2d(),
                 dw_activation=lambda_relu(), pw_activation=lambda_relu()):
        super().__init__()
        self.dw_conv = dwconv_block(in_channels=in_channels, out_channels=in_channels, kernel_size=kernel_size,
                                    stride=stride, padding=padding, dilation=dilation, bias=bias,
                                    normalization=dw_normalization, activation=dw_activation)
        self.pw_conv = conv1x1_block(in_channels=in_channels, out_channels=out_channels, bias=bias,
                                     normalization=pw_normalization, activation=pw_activation)


def dwsconv3x3_block(stride=1, padding=1, **kwargs):
    return DwsConvBlock(kernel_size=3, stride=stride, padding=padding, **kwargs)


# ---- squeeze-and-excitation (att.py:38-105) -----------------------------------------------------------------------
class SEBlock(B200Module):
    """x * out_act(W2 mid_act(W1 mean_HW(x) + b1) + b2); attribute names as att.py:59-92."""

    def __init__(self, channels, reduction=16, mid_channels=None, round_mid=False, use_conv=True,
                 mid_activation=lambda_relu(), out_activation=lambda_sigmoid()):
        super().__init__()
        self.use_conv = use_conv
        if mid_channels is None:
            mid_channels = channels // reduction if not round_mid else round_channels(float(channels) / reduction)
        self.pool = nn.AdaptiveAvgPool2d(output_size=1)
        if use_conv:
            self.conv1 = conv1x1(in_channels=channels, out_channels=mid_channels, bias=True)
        else:
            self.fc1 = nn.Linear(in_features=channels, out_features=mid_channels)
        self.activ = create_activation_layer(mid_activation)
        if use_conv:
            self.conv2 = conv1x1(in_channels=mid_channels, out_channels=channels, bias=True)
        else:
            self.fc2 = nn.Linear(in_features=mid_channels, out_features=channels)
        self.sigmoid = create_activation_layer(out_activation)
