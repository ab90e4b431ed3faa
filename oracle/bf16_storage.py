"""TEST INFRASTRUCTURE (see oracle/__init__.py): the oracle re-evaluated with the B200 bf16 tier's STORAGE format.

`oracle_forward_bf16_storage(net, x)` walks the module tree exactly like `oracle_forward` (oracle/ref_forward.py, which
restates the reference's forwards) but gives every ConvBlock / SEBlock the arithmetic contract of the bf16 tier
(DESIGN.md section 2): activations stored as bf16, BatchNorm folded into bf16 weights (conv.py:250-259 + norm.py:34-50),
fp32 accumulation, fp32 bias and activation, one rounding to bf16 per block; SE gates in fp32.  It answers "how far can
ANY implementation with this storage format be from the fp32 reference on this network?" - the error floor the bf16
parity tests compare against on networks that are ill-conditioned at random init (SURVEY hard part 3).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ref_forward as R


_STORAGE = [torch.bfloat16]   # the 16-bit element type being emulated (bf16 tier or fp16 tier)


def _rb(t: torch.Tensor) -> torch.Tensor:
    return t.to(_STORAGE[0]).float()


def _conv_block_bf16(m, x):
    """ConvBlock.forward (conv.py:278-286) with folded BN and bf16 operands / fp32 accumulate."""
    if getattr(m, "use_pad", False):
        x = F.pad(x, m.pad.padding)
    w, b = m.conv.weight, m.conv.bias
    if m.normalize:
        s = m.bn.weight / torch.sqrt(m.bn.running_var + m.bn.eps)
        w = w * s.view(-1, 1, 1, 1)
        b = (0 if b is None else b * s) + m.bn.bias - m.bn.running_mean * s
    y = F.conv2d(_rb(x), _rb(w), b, m.conv.stride, m.conv.padding, m.conv.dilation, m.conv.groups)
    if m.activate:
        y = R._activation(m.activ, y)
    return _rb(y)


def _se_block_bf16(m, x):
    """SEBlock.forward (att.py:94-105) on a bf16 map with fp32 gates."""
    return _rb(_SE(m, _rb(x)))


_CONV, _SE = R.conv_block, R.se_block


@torch.no_grad()
def oracle_forward_16bit_storage(m, x, storage=torch.bfloat16, **kw):
    """The oracle with the arithmetic contract of a 16-bit tier: `storage` = torch.bfloat16 (PCV_BF16) or torch.float16
    (PCV_F16) activations and BN-folded weights, fp32 accumulate / bias / activation, one rounding per block."""
    saved = (R.conv_block, R.se_block, R._BY_NAME["ConvBlock"], R._BY_NAME["SEBlock"], _STORAGE[0])
    R.conv_block, R.se_block = _conv_block_bf16, _se_block_bf16
    R._BY_NAME["ConvBlock"], R._BY_NAME["SEBlock"] = _conv_block_bf16, _se_block_bf16
    _STORAGE[0] = storage
    try:
        return R.oracle_forward(m, x, **kw)
    finally:
        R.conv_block, R.se_block, R._BY_NAME["ConvBlock"], R._BY_NAME["SEBlock"], _STORAGE[0] = saved


def oracle_forward_bf16_storage(m, x, **kw):
    return oracle_forward_16bit_storage(m, x, storage=torch.bfloat16, **kw)
