"""Eager, op-level access to the C ABI on torch CUDA tensors (NHWC activations).

Thin argument marshalling only: each function is exactly one entry point of include/pcv_b200.h, launched on the
current CUDA stream.  Used by the parity tests to exercise every kernel in isolation against the oracle.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import BF16, F16, F32, ConvDesc


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float16:
        return F16
    raise TypeError(f"unsupported tensor dtype {t.dtype}")


def _tdt(code: int):
    return torch.float32 if code == F32 else (torch.float16 if code == F16 else torch.bfloat16)


def _code(dtype) -> int:
    return F32 if dtype == torch.float32 else (F16 if dtype == torch.float16 else BF16)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("pytorchcv_b200 ops need CUDA tensors; there is no CPU fallback")


class PackedConv:
    """BN-folded, repacked weights of one ConvBlock for one tier."""

    def __init__(self, desc: ConvDesc, dtype: int, w: torch.Tensor, bias: torch.Tensor):
        self.desc, self.dtype, self.w, self.bias = desc, dtype, w, bias


def make_desc(N, H, W, Cin, Cout, k, stride=1, pad=0, dil=1, groups=1, act=0, in_pitch=0, out_pitch=0, res_pitch=0,
              flags=0) -> ConvDesc:
    kh, kw = (k, k) if isinstance(k, int) else k
    return ConvDesc(N=N, H=H, W=W, Cin=Cin, Cout=Cout, kh=kh, kw=kw, stride=stride, pad=pad, dil=dil, groups=groups,
                    act=act, in_pitch=in_pitch, out_pitch=out_pitch, res_pitch=res_pitch, flags=flags)


def pack_conv(desc: ConvDesc, dtype: int, weight: torch.Tensor, conv_bias=None, bn=None, eps: float = 1e-5) -> PackedConv:
    """weight: fp32 [Cout, Cin/g, kh, kw]; bn: (gamma, beta, mean, var) fp32 [Cout] or None."""
    _need_cuda(weight)
    f = lambda t: None if t is None else t.detach().to(weight.device, torch.float32).contiguous()
    weight, conv_bias = f(weight), f(conv_bias)
    g, b, m, v = (f(t) for t in bn) if bn is not None else (None, None, None, None)
    wb, bb = C.c_size_t(), C.c_size_t()
    _lib.call("pcv_conv_packed_bytes", C.byref(desc), dtype, C.byref(wb), C.byref(bb))
    wp = torch.zeros(wb.value + 16, dtype=torch.uint8, device=weight.device)
    bp = torch.zeros(bb.value // 4 + 4, dtype=torch.float32, device=weight.device)
    p = lambda t: None if t is None else t.data_ptr()
    _lib.call("pcv_pack_conv_weights", C.byref(desc), dtype, p(weight), p(conv_bias), p(g), p(b), p(m), p(v),
              float(eps), wp.data_ptr(), bp.data_ptr(), _stream())
    torch.cuda.current_stream().synchronize()  # the fp32 sources may be temporaries
    return PackedConv(desc, dtype, wp, bp)


def conv2d(x: torch.Tensor, packed: PackedConv, residual: torch.Tensor | None = None,
           out: torch.Tensor | None = None) -> torch.Tensor:
    """x: NHWC [N,H,W,in_pitch]; returns NHWC [N,Ho,Wo,out_pitch] (fp32 when the desc has CONV_OUT_F32)."""
    _need_cuda(x, residual)
    d = packed.desc
    Ho = (d.H + 2 * d.pad - d.dil * (d.kh - 1) - 1) // d.stride + 1
    Wo = (d.W + 2 * d.pad - d.dil * (d.kw - 1) - 1) // d.stride + 1
    odt = torch.float32 if (d.flags & _lib.CONV_OUT_F32) else _tdt(packed.dtype)
    if out is None:
        out = torch.zeros((d.N, Ho, Wo, d.out_pitch or d.Cout), dtype=odt, device=x.device)
    ws = C.c_size_t()
    _lib.call("pcv_conv_workspace_bytes", C.byref(d), packed.dtype, C.byref(ws))
    scratch = torch.empty(ws.value + 16, dtype=torch.uint8, device=x.device) if ws.value else None
    _lib.call("pcv_conv2d_bias_act_ws", None, C.byref(d), packed.dtype, x.data_ptr(), packed.w.data_ptr(),
              packed.bias.data_ptr(), residual.data_ptr() if residual is not None else None, out.data_ptr(),
              scratch.data_ptr() if scratch is not None else None, _stream())
    if scratch is not None:
        scratch.record_stream(torch.cuda.current_stream())
    return out


def maxpool2d(x: torch.Tensor, k: int, stride: int, pad: int) -> torch.Tensor:
    _need_cuda(x)
    N, H, W, Cc = x.shape
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    out = torch.empty((N, Ho, Wo, Cc), dtype=x.dtype, device=x.device)
    _lib.call("pcv_maxpool2d", None, _dt(x), N, H, W, Cc, k, stride, pad, x.data_ptr(), Cc, out.data_ptr(), Cc,
              _stream())
    return out


def global_avgpool(x: torch.Tensor, out_dtype=None) -> torch.Tensor:
    _need_cuda(x)
    N, H, W, Cc = x.shape
    odt = x.dtype if out_dtype is None else out_dtype
    out = torch.empty((N, Cc), dtype=odt, device=x.device)
    _lib.call("pcv_global_avgpool", None, _dt(x), N, H * W, Cc, x.data_ptr(), Cc, out.data_ptr(),
              _code(odt), _stream())
    return out


def adaptive_avgpool(x: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """nn.AdaptiveAvgPool2d((out_h, out_w)) on an NHWC tensor (pspnet.py:71)."""
    _need_cuda(x)
    N, H, W, Cc = x.shape
    out = torch.empty((N, out_h, out_w, Cc), dtype=x.dtype, device=x.device)
    _lib.call("pcv_adaptive_avgpool", None, _dt(x), N, H, W, Cc, x.data_ptr(), Cc, out_h, out_w, out.data_ptr(), _stream())
    return out


def se_excite(pooled: torch.Tensor, w1, b1, w2, b2, mid_act=_lib.ACT_RELU, out_act=_lib.ACT_SIGMOID) -> torch.Tensor:
    _need_cuda(pooled, w1, w2)
    N, Cc = pooled.shape
    cmid = w1.shape[0]
    f = lambda t: None if t is None else t.detach().to(pooled.device, torch.float32).contiguous()
    w1, b1, w2, b2 = f(w1), f(b1), f(w2), f(b2)
    buf = torch.empty(N * (Cc + cmid), dtype=torch.float32, device=pooled.device)
    p = lambda t: None if t is None else t.data_ptr()
    _lib.call("pcv_se_excite", None, N, Cc, cmid, pooled.data_ptr(), p(w1), p(b1), p(w2), p(b2), mid_act, out_act,
              buf.data_ptr(), _stream())
    torch.cuda.current_stream().synchronize()
    return buf[:N * Cc].view(N, Cc)


def se_scale_add_act(x: torch.Tensor, gate: torch.Tensor, identity=None, act=_lib.ACT_NONE) -> torch.Tensor:
    _need_cuda(x, gate, identity)
    N, H, W, Cc = x.shape
    out = torch.empty_like(x)
    _lib.call("pcv_se_scale_add_act", None, _dt(x), N, H * W, Cc, x.data_ptr(), gate.data_ptr(),
              identity.data_ptr() if identity is not None else None, act, out.data_ptr(), _stream())
    return out


def add_act(a: torch.Tensor, b: torch.Tensor, act=_lib.ACT_NONE) -> torch.Tensor:
    _need_cuda(a, b)
    out = torch.empty_like(a)
    _lib.call("pcv_add_act", None, _dt(a), a.numel(), a.data_ptr(), b.data_ptr(), act, out.data_ptr(), _stream())
    return out


def nchw_to_nhwc(x: torch.Tensor, dtype=torch.bfloat16, c_pitch: int = 0) -> torch.Tensor:
    _need_cuda(x)
    x = x.float().contiguous()
    N, Cc, H, W = x.shape
    pitch = c_pitch or (Cc + 7) // 8 * 8
    out = torch.empty((N, H, W, pitch), dtype=dtype, device=x.device)
    _lib.call("pcv_nchw_f32_to_nhwc", None, _code(dtype), N, Cc, H, W, x.data_ptr(),
              out.data_ptr(), pitch, _stream())
    return out


def nhwc_to_nchw(x: torch.Tensor, channels: int | None = None) -> torch.Tensor:
    _need_cuda(x)
    N, H, W, pitch = x.shape
    Cc = channels or pitch
    out = torch.empty((N, Cc, H, W), dtype=torch.float32, device=x.device)
    _lib.call("pcv_nhwc_to_nchw_f32", None, _dt(x), N, Cc, H, W, x.data_ptr(), pitch, out.data_ptr(), _stream())
    return out


def bilinear_upsample_ac(x: torch.Tensor, Hout: int, Wout: int, channels: int | None = None,
                         nchw_f32: bool = False) -> torch.Tensor:
    _need_cuda(x)
    N, H, W, pitch = x.shape
    Cc = channels or pitch
    if nchw_f32:
        out = torch.empty((N, Cc, Hout, Wout), dtype=torch.float32, device=x.device)
    else:
        out = torch.empty((N, Hout, Wout, Cc), dtype=x.dtype, device=x.device)
    _lib.call("pcv_bilinear_upsample_ac", None, _dt(x), N, H, W, Cc, x.data_ptr(), pitch, Hout, Wout, out.data_ptr(),
              Cc, 1 if nchw_f32 else 0, _stream())
    return out
