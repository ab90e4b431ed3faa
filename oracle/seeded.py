"""ORACLE — test infrastructure only.  Order-independent, name-keyed deterministic weights and inputs.

`torch.manual_seed(s); get_model(...)` reproduces the reference's own random init only if two implementations
construct their modules in the same order.  For parity tests we want something stronger: every tensor of the
state_dict is filled from a generator seeded by (seed, crc32(key)), so the reference module tree and the mirror
module tree receive bit-identical values key by key regardless of construction order.

`randomize_bn=True` also draws BatchNorm gamma in U(0.5,1.5), beta / running_mean in U(-0.5,0.5), running_var in
U(0.5,1.5) and biases in U(-0.5,0.5): with the reference's default init every BatchNorm is the identity and every
bias is 0, which would leave BN-folding and bias bugs invisible (hard part 1).
"""
from __future__ import annotations

import zlib

import torch


def _gen(seed: int, key: str) -> torch.Generator:
    return torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 63 - 1))


@torch.no_grad()
def seeded_init(net: torch.nn.Module, seed: int = 0, randomize_bn: bool = True) -> torch.nn.Module:
    """Fill every parameter / buffer of `net` in place (CPU or CUDA) from name-keyed generators."""
    for key, t in net.state_dict().items():
        g = _gen(seed, key)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            continue
        if t.dim() >= 2:  # conv / linear weight: kaiming-uniform bound sqrt(6 / fan_in), as resnet.py:326-331
            fan_in = t[0].numel()
            bound = (6.0 / fan_in) ** 0.5
            val = (torch.rand(t.shape, generator=g) * 2 - 1) * bound
        elif leaf == "running_var":
            val = torch.rand(t.shape, generator=g) + 0.5 if randomize_bn else torch.ones(t.shape)
        elif leaf == "running_mean":
            val = torch.rand(t.shape, generator=g) - 0.5 if randomize_bn else torch.zeros(t.shape)
        elif leaf == "weight":  # BatchNorm gamma (1-D weight)
            val = torch.rand(t.shape, generator=g) + 0.5 if randomize_bn else torch.ones(t.shape)
        elif leaf == "bias":
            val = torch.rand(t.shape, generator=g) - 0.5 if randomize_bn else torch.zeros(t.shape)
        else:
            continue
        t.copy_(val.to(t.dtype))
    return net


def seeded_input(shape, seed: int = 1234) -> torch.Tensor:
    """The synthetic batch of SURVEY 8d: torch.randn(N,3,H,W) from Generator(seed), fp32 NCHW on CPU."""
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))
