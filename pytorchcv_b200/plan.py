"""Compile a pytorchcv module tree into a flat plan of fused sm_100a kernels and run it.

The reference executes ~4 eager torch ops per block from Python (ConvBlock.forward, pytorchcv/models/common/conv.py:
278-286).  Here `compile_module` walks the module tree ONCE, pattern-matching on the reference's class names and
attributes (so it accepts both this package's mirror modules and real `pytorchcv` modules), folds BatchNorm into
packed weights, lays activations out as NHWC in one arena with liveness-based reuse, and records one fused C-ABI op
per block into a `pcv_plan`.  A forward is then: one ingest kernel (NCHW fp32 -> NHWC) + `pcv_plan_run`.

Nothing here computes on the CPU or through torch ops: every op goes through libpcv_b200.so, and unsupported module
patterns raise (NotImplementedError / ValueError) at compile time.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from dataclasses import dataclass, field
from typing import Any, Callable

import torch
import torch.nn as nn

from . import _lib
from ._lib import (ACT_NONE, ACT_RELU, ACT_RELU6, ACT_SIGMOID, ACT_SWISH, ACT_HSWISH, ACT_HSIGMOID, ACT_LEAKY_RELU, ACT_CLAMP01,
                   BF16, F16, F32, ConvDesc)

# fp32 tier: dense / grouped convs on the tensor cores as a 3-way bf16 split (PCV_F32_SPLIT=0: the CUDA-core kernel instead)
_F32_SPLIT = os.environ.get("PCV_F32_SPLIT", "1") != "0"
# cross-layer fusion of bottleneck tails (pcv_bottleneck_tail): correct, but measured no faster than the two kernels it
# replaces (shared-memory-port bound, csrc/conv_igemm3x.cu) - off unless asked for
_FUSE_TAIL = [os.environ.get("PCV_FUSE_TAIL", "0") == "1"]


def set_fuse_tail(enabled: bool) -> None:
    """Record ResBottleneck tails (3x3 -> 1x1 + add + ReLU) as one fused kernel where the shapes allow it."""
    _FUSE_TAIL[0] = bool(enabled)


# SE units: squeeze moved upstream of the unit's last 1x1 conv by linearity (PCV_SE_FOLD=0: pool the conv's output instead)
_SE_FOLD = os.environ.get("PCV_SE_FOLD", "1") != "0"
# depthwise -> pointwise pairs as one fused kernel (pcv_dw_pw_fused); PCV_FUSE_DWPW=0 records the two convolutions
_FUSE_DWPW = [os.environ.get("PCV_FUSE_DWPW", "1") != "0"]


def set_fuse_dwpw(enabled: bool) -> None:
    _FUSE_DWPW[0] = bool(enabled)


# 1x1 expansion -> depthwise -> pointwise triples as one fused kernel (pcv_exp_dw_pw_fused); PCV_FUSE_XDWPW=0 records the
# expansion and the dw -> pw pair
_FUSE_XDWPW = [os.environ.get("PCV_FUSE_XDWPW", "1") != "0"]
# Measured (MobileNetV2 bs256, fp16): the triple kernel is bound by its CUDA-core work (stencil + conversion of the expanded halo),
# not by HBM, so it pays where the expanded tensor is largest relative to that work - stride 2, whose two-kernel plan moves 4
# expanded pixels per output pixel (16->96->24 @112: 375 -> 327 us).  Stride-1 triples are at parity with expansion +
# pcv_dw_pw_fused and are recorded only on request (PCV_XDWPW_S1=1 / set_fuse_xdwpw(True, stride1=True)).
_XDWPW_S1 = [os.environ.get("PCV_XDWPW_S1", "0") == "1"]


def set_fuse_xdwpw(enabled: bool, stride1: bool | None = None) -> None:
    _FUSE_XDWPW[0] = bool(enabled)
    if stride1 is not None:
        _XDWPW_S1[0] = bool(stride1)


def _exp_dw_pw(b, expb, dwb, pwb, x, residual=None, post_act=None):
    """expansion ConvBlock -> dw ConvBlock -> pw ConvBlock [+ residual]: one kernel where the triple is in its domain, else the
    expansion followed by the dw -> pw pair."""
    fused = b.exp_dw_pw(x, expb, dwb, pwb, residual, post_act)
    if fused is not None:
        return fused
    return _dw_then_pw(b, dwb, pwb, lower(b, expb, x), residual=residual, post_act=post_act)


def _dw_then_pw(b, dwb, pwb, x, residual=None, post_act=None, **kw):
    """dw ConvBlock -> pw ConvBlock [+ residual, post_act]: the fused kernel where it applies, else the two convolutions."""
    if not kw:
        fused = b.dw_pw(x, dwb, pwb, residual, post_act)
        if fused is not None:
            return fused
    return lower(b, pwb, lower(b, dwb, x), residual=residual, post_act=post_act, **kw)


_ALIGN = 1024  # arena / weight blob alignment (TMA needs 16 B; 1 KiB keeps every tensor sector- and line-aligned)


def _esize(dtype: int) -> int:
    return 4 if dtype == F32 else 2


def _rup(a: int, b: int) -> int:
    return (a + b - 1) // b * b


_DTYPE_NAMES = {"bf16": BF16, "bfloat16": BF16, "fp16": F16, "f16": F16, "half": F16, "float16": F16,
                "fp32": F32, "f32": F32, "float32": F32,
                torch.bfloat16: BF16, torch.float16: F16, torch.float32: F32}


def dtype_code(dtype) -> int:
    """Tier code (pcv_dtype) of "bf16" | "fp16" | "fp32", a torch dtype, or a code."""
    if isinstance(dtype, int) and not isinstance(dtype, bool) and dtype in (BF16, F32, F16):
        return dtype
    try:
        return _DTYPE_NAMES[dtype]
    except (KeyError, TypeError):
        raise ValueError(f"unsupported dtype {dtype!r}: the path has two 16-bit storage tiers (bf16, fp16) and an "
                         "fp32 tier") from None


def _is16(dtype: int) -> bool:
    return dtype in (BF16, F16)


# ---------------------------------------------------------------------------------------------------------------
# symbolic tensors
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class Buf:
    nbytes: int
    first: int            # op index that defines it (-1: network input)
    last: int = -1        # last op index that reads it
    pinned: bool = False  # network input / output: never reused
    offset: int = -1


@dataclass
class TRef:
    """A logical NHWC activation: N x H x W pixels, C channels at `pitch` elements per pixel, inside `buf`."""
    N: int
    H: int
    W: int
    C: int
    pitch: int
    dtype: int
    buf: Buf
    ch_off: int = 0
    layout: str = "nhwc"   # "nhwc" | "nchw" (fp32 network outputs written by the bilinear kernel)
    flat: bool = False     # return as [N, C] (the reference's x.view(N, -1))
    tail: Any = None       # network-edge op run per call AFTER the plan, writing a fresh caller-owned tensor:
                           # tail(ptr, out_ptr, stream); such a tensor has no arena storage and cannot feed another op
    cmap: Any = None       # "virtual channel padding": storage index of each REAL channel when the C storage channels hold
                           # fewer real ones (a torch.cat of parts whose widths are not multiples of 8, ghostnet.py:57-60: every
                           # part is padded to 8 so that all slices stay 16-byte aligned).  Padding channels hold exact zeros;
                           # consumers scatter their weights' input channels accordingly (zero weights on the padding).

    @property
    def creal(self) -> int:
        return self.C if self.cmap is None else len(self.cmap)

    @property
    def byte_off(self) -> int:
        return self.ch_off * _esize(self.dtype)


# ---------------------------------------------------------------------------------------------------------------
# activation / norm classification (activ.py:188-222, norm.py:95-115)
# ---------------------------------------------------------------------------------------------------------------
def act_code(m: nn.Module | None) -> int:
    if m is None or isinstance(m, nn.Identity):
        return ACT_NONE
    if isinstance(m, nn.ReLU6):
        return ACT_RELU6
    if isinstance(m, nn.ReLU):
        return ACT_RELU
    if isinstance(m, nn.Sigmoid):
        return ACT_SIGMOID
    name = type(m).__name__
    if name == "Swish" or isinstance(m, nn.SiLU):
        return ACT_SWISH
    if name == "HSwish" or isinstance(m, nn.Hardswish):
        return ACT_HSWISH
    if name == "HSigmoid" or isinstance(m, nn.Hardsigmoid):
        return ACT_HSIGMOID
    if name == "GhostHSigmoid":
        return ACT_CLAMP01               # clamp(x, 0, 1) (ghostnet.py:18-24): SE gates only
    if isinstance(m, nn.LeakyReLU):
        return ACT_LEAKY_RELU            # slope: act_param(m); fused into dense / grouped conv epilogues
    raise NotImplementedError(f"activation {name} is outside the B200 eval path (SURVEY 8a8)")


def act_param(m: nn.Module | None) -> float:
    """The scalar an activation code carries: nn.LeakyReLU's negative slope (activ.py:101-120)."""
    return float(m.negative_slope) if isinstance(m, nn.LeakyReLU) else 0.0


def _standalone_act(m: nn.Module | None, conv: nn.Conv2d | None = None) -> bool:
    """Activations that do not ride on the conv's epilogue and run as one pcv_channel_affine_act pass behind it: nn.PReLU
    (per-channel slopes, activ.py:84-98) and a LeakyReLU behind a depthwise conv."""
    if isinstance(m, nn.PReLU):
        return True
    return isinstance(m, nn.LeakyReLU) and conv is not None and conv.groups > 1 and conv.groups == conv.in_channels


def _bn_scale_shift(bn: nn.BatchNorm2d):
    """Eval BatchNorm as y = x * scale + shift (norm.py:34-50), computed in float64 on the host."""
    _check_bn(bn)
    with torch.no_grad():
        var, mean = bn.running_var.detach().double(), bn.running_mean.detach().double()
        g = bn.weight.detach().double() if bn.weight is not None else torch.ones_like(var)
        be = bn.bias.detach().double() if bn.bias is not None else torch.zeros_like(var)
        sc = g / torch.sqrt(var + bn.eps)
        return sc.float().contiguous(), (be - mean * sc).float().contiguous()


def _check_bn(bn: nn.Module) -> None:
    if not isinstance(bn, nn.BatchNorm2d):
        raise NotImplementedError(f"normalization {type(bn).__name__} cannot be folded (needs per-sample statistics)")
    if bn.training:
        raise RuntimeError("pytorchcv_b200 is an eval-mode path: call net.eval() first (BatchNorm uses running stats)")
    if bn.running_mean is None or bn.running_var is None:
        raise NotImplementedError("BatchNorm2d without running statistics cannot be folded")


def _one(v) -> int:
    if isinstance(v, (tuple, list)):
        if len(v) != 2 or v[0] != v[1]:
            raise NotImplementedError(f"only square kernels / symmetric stride, padding, dilation are supported, got {v}")
        return int(v[0])
    return int(v)


# ---------------------------------------------------------------------------------------------------------------
# builder: records symbolic ops, then materialises them
# ---------------------------------------------------------------------------------------------------------------
class _NoS2dStem(Exception):
    """The network input feeds more than the stem convolution: recompile with the generic NHWC ingest."""


class _NoStemPool(Exception):
    """The stem's conv map is read by something besides the fused max pool: recompile without that fusion."""


class Builder:
    def __init__(self, dtype: int, device: torch.device, allow_s2d_stem: bool = True, allow_stem_pool: bool = True):
        self.dtype = dtype
        self.device = device
        self.allow_s2d_stem = allow_s2d_stem
        self.allow_stem_pool = allow_stem_pool
        self.input_tref: TRef | None = None   # NHWC (channel-padded) image, set by CompiledModule
        self.image_channels = 0
        self.stem: dict | None = None         # {"k": kernel, "tref": s2d tensor} once the s2d stem is chosen
        self.flops_override: dict[int, float] = {}   # op index -> algorithmic FLOPs where the kernel's GEMM view pads K
        self.ops: list[Callable[[Any, Callable[[TRef], int]], None]] = []
        self.bufs: list[Buf] = []
        self.weight_jobs: list[tuple] = []   # (kind, payload) resolved in materialise()
        self.weight_bytes = 0

    # -- buffers ------------------------------------------------------------------------------------------------
    def new(self, N, H, W, C, dtype=None, pitch=None, layout="nhwc", extra_bytes=0) -> TRef:
        dtype = self.dtype if dtype is None else dtype
        pitch = C if pitch is None else pitch
        buf = Buf(nbytes=N * H * W * pitch * _esize(dtype) + extra_bytes, first=len(self.ops))
        self.bufs.append(buf)
        return TRef(N, H, W, C, pitch, dtype, buf, 0, layout)

    @staticmethod
    def view(t: TRef, ch_off: int, C: int) -> TRef:
        return TRef(t.N, t.H, t.W, C, t.pitch, t.dtype, t.buf, t.ch_off + ch_off, t.layout)

    def _use(self, *trefs: TRef | None) -> int:
        idx = len(self.ops)
        for t in trefs:
            if t is not None:
                if t.tail is not None:
                    raise NotImplementedError("an fp32 NCHW network output cannot feed another op of the plan")
                if self.stem is not None and self.input_tref is not None and t.buf is self.input_tref.buf:
                    raise _NoS2dStem()
                if self.stem is not None and t.buf is self.stem.get("fused_conv_buf"):
                    raise _NoStemPool()
                t.buf.last = max(t.buf.last, idx)
        return idx

    def _s2d_stem(self, x: TRef, conv, k: int, stride: int, pad: int, dil: int) -> bool:
        """k x k stride-2 conv on the <=4-channel network input -> space-to-depth stem (include/pcv_b200.h)."""
        return (self.allow_s2d_stem and _is16(self.dtype) and x is self.input_tref and self.stem is None
                and x.buf.last < 0 and conv.groups == 1 and conv.in_channels == self.image_channels <= 4
                and k in (3, 5, 7) and conv.kernel_size[0] == conv.kernel_size[1] and stride == 2 and pad == k // 2
                and dil == 1 and x.H % 2 == 0 and x.W % 2 == 0)

    def _conv_s2d(self, x: TRef, conv, bn, act: int, k: int) -> TRef:
        rows, cols, cin_eq, taps = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _lib.call("pcv_stem_s2d_dims", conv.in_channels, x.H, x.W, k, C.byref(rows), C.byref(cols), C.byref(cin_eq),
                  C.byref(taps))
        xs = self.new(x.N, rows.value, cols.value, 16)
        xs.buf.first, xs.buf.pinned = -1, True
        x.buf.nbytes = 0                      # the NHWC copy of the image is never materialised
        self.stem = {"k": k, "tref": xs}
        Ho, Wo, cout = x.H // 2, x.W // 2, conv.out_channels
        out = self.new(x.N, Ho, Wo, cout)
        d = ConvDesc(N=x.N, H=rows.value, W=Wo, Cin=cin_eq.value, Cout=cout, kh=taps.value, kw=1, stride=1, pad=0,
                     dil=1, groups=1, act=act, in_pitch=16, out_pitch=out.pitch, res_pitch=0,
                     flags=_lib.CONV_IN_OVERLAP, in_row_pitch=cols.value * 16)
        wb, bb = C.c_size_t(), C.c_size_t()
        _lib.call("pcv_conv_packed_bytes", C.byref(d), self.dtype, C.byref(wb), C.byref(bb))
        w_off, b_off = self._wblob(wb.value), self._wblob(bb.value)
        self.weight_jobs.append(("conv_s2d", (d, conv, bn, k, w_off, b_off)))
        idx = len(self.ops)
        xs.buf.last = out.buf.last = idx
        # FLOPs = 2 * MACs of the reference's k x k conv on C channels (the s2d GEMM computes (k+1)^2 * 4C
        # products per output, zero weights included - that padding is not algorithmic work)
        self.flops_override[idx] = 2.0 * x.N * Ho * Wo * cout * conv.in_channels * k * k
        dtype = self.dtype
        target = [out]   # maxpool() may retarget the op at the pooled map (PCV_CONV_POOL3S2)
        self.stem.update(out=out, idx=idx, desc=d, target=target,
                         pool_ok=bool(_lib.load().pcv_stem_s2d_pool_ok(conv.in_channels, x.H, x.W, k, cout)))

        def emit(plan, ptr, wptr):
            _lib.call("pcv_conv2d_bias_act", plan, C.byref(d), dtype, ptr(xs), wptr(w_off), wptr(b_off), None,
                      ptr(target[0]), None)
        self.ops.append(emit)
        return out

    def _wblob(self, nbytes: int) -> int:
        off = self.weight_bytes
        self.weight_bytes += _rup(max(nbytes, 16), 256)
        return off

    # -- ops ----------------------------------------------------------------------------------------------------
    def conv(self, x: TRef, conv: nn.Conv2d, bn: nn.Module | None = None, act: int = ACT_NONE,
             residual: TRef | None = None, out: TRef | None = None, out_f32: bool = False, flags: int = 0,
             pad_lrtb: tuple | None = None, act_a: float = 0.0, out_cmap: list | None = None,
             gate: TRef | None = None) -> TRef | None:
        """One fused ConvBlock: conv + folded BN + optional residual + activation.  `pad_lrtb` = (left, right, top, bottom)
        replaces the conv's own padding (ZeroPad2d / tf_mode): symmetric amounts ride on the kernel's padding, asymmetric
        ones are materialised by one zero-pad pass."""
        if conv.padding_mode != "zeros" or isinstance(conv.padding, str):
            raise NotImplementedError("only zero padding with integer sizes is supported")
        kh, kw = conv.kernel_size
        k_stride, k_pad, k_dil = _one(conv.stride), _one(conv.padding), _one(conv.dilation)
        if pad_lrtb is not None:
            if k_pad != 0:
                raise NotImplementedError("explicit padding in front of a convolution that pads itself")
            pl, pr, pt, pb = (int(v) for v in pad_lrtb)
            if pl == pr == pt == pb:
                k_pad = pl
            else:
                x = self.pad(x, pl, pr, pt, pb)
        cin, cout, groups = conv.in_channels, conv.out_channels, conv.groups
        if gate is not None and (x.cmap is not None or out_cmap is not None or out is not None or not _is16(self.dtype)):
            return None
        # virtual channel padding (TRef.cmap): the input's / output's real channels sit at given storage positions
        if x.cmap is not None or out_cmap is not None or (out is not None and out.C != cout):
            return self._conv_mapped(x, conv, bn, act, act_a, residual, out, out_cmap, k_stride, k_pad, k_dil, flags)
        if (residual is None and out is None and not out_f32 and act != ACT_LEAKY_RELU
                and self._s2d_stem(x, conv, kh, k_stride, k_pad, k_dil)):
            if bn is not None:
                _check_bn(bn)
            return self._conv_s2d(x, conv, bn, act, kh)
        pad_cin = 0
        if cin != x.C:
            # the ingest pads the image's channels with zeros up to a multiple of 8 (TMA needs 16-byte strides):
            # widen the weights with zero input channels to match
            if groups != 1 or x.C < cin:
                raise ValueError(f"conv expects {cin} input channels, tensor has {x.C}")
            pad_cin = x.C - cin
        if bn is not None:
            _check_bn(bn)
        Ho = (x.H + 2 * k_pad - k_dil * (kh - 1) - 1) // k_stride + 1
        Wo = (x.W + 2 * k_pad - k_dil * (kw - 1) - 1) // k_stride + 1
        if Ho <= 0 or Wo <= 0:
            raise ValueError(f"convolution output is empty for input {x.H}x{x.W}")
        odt = F32 if out_f32 else self.dtype
        if out is None:
            out = self.new(x.N, Ho, Wo, cout, dtype=odt)
        if (out.N, out.H, out.W, out.C, out.dtype) != (x.N, Ho, Wo, cout, odt):
            raise ValueError("conv output view has the wrong shape")
        if residual is not None and (residual.N, residual.H, residual.W, residual.C) != (x.N, Ho, Wo, cout):
            raise ValueError("residual shape does not match the conv output")
        if self.dtype == F32 and _F32_SPLIT and groups != cin:
            # the fp32 tier's dense / grouped convs run on the tensor cores as a 3-way bf16 split (include/pcv_b200.h
            # PCV_CONV_F32_SPLIT); the library ignores the flag (CUDA-core kernel) for shapes the tcgen05 route cannot take
            flags |= _lib.CONV_F32_SPLIT
        if gate is not None:
            # SE scale + identity + activation in this conv's epilogue (include/pcv_b200.h PCV_CONV_SE_GATE); None when the layer
            # is outside that kernel's domain - the caller then records the conv and the SE scale pass separately
            flags |= _lib.CONV_SE_GATE
        d = ConvDesc(N=x.N, H=x.H, W=x.W, Cin=x.C, Cout=cout, kh=kh, kw=kw, stride=k_stride, pad=k_pad, dil=k_dil,
                     groups=groups, act=act, in_pitch=x.pitch, out_pitch=out.pitch,
                     res_pitch=residual.pitch if residual is not None else 0,
                     flags=flags | (_lib.CONV_OUT_F32 if (out_f32 and _is16(self.dtype)) else 0), act_param=act_a)
        if gate is not None:
            if (gate.C != cout or gate.N != x.N or gate.dtype != F32 or gate.cmap is not None
                    or not _lib.load().pcv_conv_se_gate_ok(C.byref(d), self.dtype)):
                self.bufs.remove(out.buf)   # nothing was recorded: drop the output buffer allocated above
                return None
        wb, bb, ws = C.c_size_t(), C.c_size_t(), C.c_size_t()
        _lib.call("pcv_conv_packed_bytes", C.byref(d), self.dtype, C.byref(wb), C.byref(bb))
        _lib.call("pcv_conv_workspace_bytes", C.byref(d), self.dtype, C.byref(ws))
        w_off, b_off = self._wblob(wb.value), self._wblob(bb.value)
        self.weight_jobs.append(("conv", (d, conv, bn, pad_cin, w_off, b_off)))
        idx = self._use(x, residual, out, gate)
        scratch = gate   # PCV_CONV_SE_GATE: the `workspace` argument carries the gate
        if ws.value:   # private to this op: lives exactly as long as op `idx`
            sbuf = Buf(nbytes=ws.value, first=idx, last=idx)
            self.bufs.append(sbuf)
            scratch = TRef(1, 1, 1, ws.value // 2, ws.value // 2, BF16, sbuf)
        dtype = self.dtype

        def emit(plan, ptr, wptr):
            _lib.call("pcv_conv2d_bias_act_ws", plan, C.byref(d), dtype, ptr(x), wptr(w_off), wptr(b_off),
                      ptr(residual) if residual is not None else None, ptr(out),
                      ptr(scratch) if scratch is not None else None, None)
        self.ops.append(emit)
        return out

    def conv_dual(self, x: TRef, cb: nn.Module, x2: TRef, cb2: nn.Module, act: int, gate: TRef | None = None) -> TRef | None:
        """act(cb(x) + cb2(x2)) for two linear 1x1 ConvBlocks as ONE GEMM over K-concatenated operands (include/pcv_b200.h
        pcv_conv1x1_dual): a bottleneck's conv3 with the unit's projection shortcut folded in - the identity tensor is never
        written.  With `gate` (an SE unit): act(cb(x) * gate + cb2(x2)) (pcv_conv1x1_dual_se).  None when the pair is outside
        that kernel's domain."""
        if not _DUAL_IDENTITY[0] or not _is16(self.dtype) or x.cmap is not None or x2.cmap is not None:
            return None
        if gate is not None and (x.H * x.W < _DUAL_GATE_MIN_HW[0] or gate.dtype != F32 or gate.cmap is not None
                                 or gate.N != x.N or gate.C != cb.conv.out_channels):
            return None
        for m in (cb, cb2):
            c = getattr(m, "conv", None)
            if (type(m).__name__ != "ConvBlock" or m.activate or getattr(m, "use_pad", False) or c is None
                    or c.kernel_size != (1, 1) or c.groups != 1 or c.padding_mode != "zeros" or isinstance(c.padding, str)):
                return None
        c1, c2 = cb.conv, cb2.conv
        try:
            s1, p1, s2, p2 = _one(c1.stride), _one(c1.padding), _one(c2.stride), _one(c2.padding)
        except NotImplementedError:
            return None
        if (s1, p1, p2) != (1, 0, 0) or c1.in_channels != x.C or c2.in_channels != x2.C or c1.out_channels != c2.out_channels:
            return None
        cout = c1.out_channels
        d = ConvDesc(N=x.N, H=x.H, W=x.W, Cin=x.C, Cout=cout, kh=1, kw=1, stride=1, pad=0, dil=1, groups=1, act=act,
                     in_pitch=x.pitch, out_pitch=cout, res_pitch=0, flags=_lib.CONV_SE_GATE if gate is not None else 0)
        d2 = ConvDesc(N=x2.N, H=x2.H, W=x2.W, Cin=x2.C, Cout=cout, kh=1, kw=1, stride=s2, pad=0, dil=1, groups=1,
                      act=ACT_NONE, in_pitch=x2.pitch, out_pitch=cout, res_pitch=0, flags=0)
        if not _lib.load().pcv_conv1x1_dual_ok(C.byref(d), C.byref(d2), self.dtype):
            return None
        for m in (cb, cb2):
            if m.normalize:
                _check_bn(m.bn)
        out = self.new(x.N, x.H, x.W, cout)
        sizes = []
        for dd in (d, d2):
            wb, bb = C.c_size_t(), C.c_size_t()
            _lib.call("pcv_conv_packed_bytes", C.byref(dd), self.dtype, C.byref(wb), C.byref(bb))
            sizes.append((wb.value, bb.value))
        w_off, b_off = self._wblob(sizes[0][0] + sizes[1][0]), self._wblob(sizes[0][1])
        b2_off = self._wblob(sizes[1][1]) if gate is not None else None   # gated: the shortcut's bias stays outside the gate
        self.weight_jobs.append(("conv_dual", (d, cb, d2, cb2, sizes, w_off, b_off, b2_off)))
        self._use(x, x2, out, gate)
        dtype = self.dtype

        def emit(plan, ptr, wptr):
            if gate is not None:
                _lib.call("pcv_conv1x1_dual_se", plan, C.byref(d), C.byref(d2), dtype, ptr(x), ptr(x2), wptr(w_off),
                          wptr(b_off), wptr(b2_off), ptr(gate), ptr(out), None)
            else:
                _lib.call("pcv_conv1x1_dual", plan, C.byref(d), C.byref(d2), dtype, ptr(x), ptr(x2), wptr(w_off),
                          wptr(b_off), ptr(out), None)
        self.ops.append(emit)
        return out

    def _conv_mapped(self, x: TRef, conv, bn, act: int, act_a: float, residual: TRef | None, out: TRef | None,
                     out_cmap: list | None, k_stride: int, k_pad: int, k_dil: int, flags: int) -> TRef:
        """A dense or depthwise conv whose input and / or output carries virtual channel padding (TRef.cmap).  The weights are
        scattered into the storage channel space - zero rows for padding outputs (with bias 0 and an identity BatchNorm, so
        they stay exact zeros through any activation with act(0) = 0), zero columns for padding inputs."""
        kh, kw = conv.kernel_size
        cin, cout, groups = conv.in_channels, conv.out_channels, conv.groups
        depthwise = groups > 1 and groups == cin and cin == cout
        if groups != 1 and not depthwise:
            raise NotImplementedError("grouped convolution on a tensor with virtual channel padding")
        if x.creal != cin:
            raise ValueError(f"conv expects {cin} input channels, tensor has {x.creal}")
        if act in (ACT_SIGMOID, ACT_HSIGMOID):
            raise NotImplementedError("act(0) != 0 on a tensor with virtual channel padding")
        in_idx = list(x.cmap) if x.cmap is not None else list(range(cin))
        if depthwise:
            out_idx, cs_out = in_idx, x.C
            if out_cmap is not None and list(out_cmap) != in_idx:
                raise NotImplementedError("a depthwise conv keeps its input's channel layout")
            if out is not None and out.C != cs_out:
                raise ValueError("depthwise output view has the wrong width")
        else:
            out_idx = list(out_cmap) if out_cmap is not None else list(range(cout))
            cs_out = out.C if out is not None else _rup(max(out_idx) + 1, 8)
        if len(out_idx) != cout or max(out_idx) >= cs_out:
            raise ValueError("output channel map does not match the convolution")
        if bn is not None:
            _check_bn(bn)
        Ho = (x.H + 2 * k_pad - k_dil * (kh - 1) - 1) // k_stride + 1
        Wo = (x.W + 2 * k_pad - k_dil * (kw - 1) - 1) // k_stride + 1
        if Ho <= 0 or Wo <= 0:
            raise ValueError(f"convolution output is empty for input {x.H}x{x.W}")
        if out is None:
            out = self.new(x.N, Ho, Wo, cs_out)
        if (out.N, out.H, out.W, out.dtype) != (x.N, Ho, Wo, self.dtype):
            raise ValueError("conv output view has the wrong shape")
        identity_map = out_idx == list(range(cout)) and cs_out == cout
        out.cmap = None if identity_map else out_idx
        if residual is not None:
            if (residual.N, residual.H, residual.W, residual.C) != (x.N, Ho, Wo, cs_out) or \
                    (list(residual.cmap) if residual.cmap is not None else list(range(cs_out))) != \
                    (out_idx if not identity_map else list(range(cs_out))):
                raise ValueError("residual layout does not match the conv output")
        if self.dtype == F32 and _F32_SPLIT and not depthwise:
            flags |= _lib.CONV_F32_SPLIT
        d = ConvDesc(N=x.N, H=x.H, W=x.W, Cin=x.C, Cout=cs_out, kh=kh, kw=kw, stride=k_stride, pad=k_pad, dil=k_dil,
                     groups=x.C if depthwise else 1, act=act, in_pitch=x.pitch, out_pitch=out.pitch,
                     res_pitch=residual.pitch if residual is not None else 0, flags=flags, act_param=act_a)
        wb, bb, ws = C.c_size_t(), C.c_size_t(), C.c_size_t()
        _lib.call("pcv_conv_packed_bytes", C.byref(d), self.dtype, C.byref(wb), C.byref(bb))
        _lib.call("pcv_conv_workspace_bytes", C.byref(d), self.dtype, C.byref(ws))
        w_off, b_off = self._wblob(wb.value), self._wblob(bb.value)
        self.weight_jobs.append(("conv_mapped", (d, conv, bn, in_idx, x.C, out_idx, cs_out, depthwise, w_off, b_off)))
        idx = self._use(x, residual, out)
        scratch = None
        if ws.value:
            sbuf = Buf(nbytes=ws.value, first=idx, last=idx)
            self.bufs.append(sbuf)
            scratch = TRef(1, 1, 1, ws.value // 2, ws.value // 2, BF16, sbuf)
        dtype = self.dtype

        def emit(plan, ptr, wptr):
            _lib.call("pcv_conv2d_bias_act_ws", plan, C.byref(d), dtype, ptr(x), wptr(w_off), wptr(b_off),
                      ptr(residual) if residual is not None else None, ptr(out),
                      ptr(scratch) if scratch is not None else None, None)
        self.ops.append(emit)
        return out

    def bottleneck_tail(self, x: TRef, cb2: nn.Module, cb3: nn.Module, residual: TRef, post_act: int) -> TRef | None:
        """conv2 (3x3 ConvBlock) -> conv3 (1x1 ConvBlock) + residual + the unit's activation as ONE fused kernel
        (include/pcv_b200.h pcv_bottleneck_tail), or None when the pair is outside that kernel's domain."""
        if not _FUSE_TAIL[0] or not _is16(self.dtype) or residual is None or post_act not in (ACT_RELU, ACT_RELU6):
            return None
        if x.cmap is not None or residual.cmap is not None:
            return None
        for cb in (cb2, cb3):
            if (type(cb).__name__ != "ConvBlock" or getattr(cb, "use_pad", False) or not cb.normalize
                    or cb.conv.padding_mode != "zeros" or isinstance(cb.conv.padding, str) or cb.conv.bias is not None):
                return None
        if not cb2.activate or act_code(cb2.activ) != ACT_RELU or cb3.activate:
            return None
        c2, c3 = cb2.conv, cb3.conv
        if c2.in_channels != x.C or c3.in_channels != c2.out_channels:
            return None
        try:
            geo = [(c.kernel_size, _one(c.stride), _one(c.padding), _one(c.dilation), c.groups) for c in (c2, c3)]
        except NotImplementedError:
            return None
        (k2, s2, p2, dl2, g2), (k3, s3, p3, dl3, g3) = geo
        d2 = ConvDesc(N=x.N, H=x.H, W=x.W, Cin=x.C, Cout=c2.out_channels, kh=k2[0], kw=k2[1], stride=s2, pad=p2, dil=dl2,
                      groups=g2, act=ACT_RELU, in_pitch=x.pitch, out_pitch=c2.out_channels, res_pitch=0, flags=0)
        d3 = ConvDesc(N=x.N, H=x.H, W=x.W, Cin=c3.in_channels, Cout=c3.out_channels, kh=k3[0], kw=k3[1], stride=s3, pad=p3,
                      dil=dl3, groups=g3, act=post_act, in_pitch=c3.in_channels, out_pitch=c3.out_channels,
                      res_pitch=residual.pitch, flags=0)
        if (residual.N, residual.H, residual.W, residual.C) != (x.N, x.H, x.W, c3.out_channels):
            return None
        if not _lib.load().pcv_bottleneck_tail_fusable(C.byref(d2), C.byref(d3), self.dtype):
            return None
        _check_bn(cb2.bn)
        _check_bn(cb3.bn)
        out = self.new(x.N, x.H, x.W, c3.out_channels)
        offs = []
        for d, cb in ((d2, cb2), (d3, cb3)):
            wb, bb = C.c_size_t(), C.c_size_t()
            _lib.call("pcv_conv_packed_bytes", C.byref(d), self.dtype, C.byref(wb), C.byref(bb))
            w_off, b_off = self._wblob(wb.value), self._wblob(bb.value)
            self.weight_jobs.append(("conv", (d, cb.conv, cb.bn, 0, w_off, b_off)))
            offs += [w_off, b_off]
        self._use(x, residual, out)
        dtype = self.dtype

        def emit(plan, ptr, wptr):
            _lib.call("pcv_bottleneck_tail", plan, C.byref(d2), C.byref(d3), dtype, ptr(x), wptr(offs[0]), wptr(offs[1]),
                      wptr(offs[2]), wptr(offs[3]), ptr(residual), ptr(out), None)
        self.ops.append(emit)
        return out

    def pad(self, x: TRef, left: int, right: int, top: int, bottom: int) -> TRef:
        """nn.ZeroPad2d((left, right, top, bottom)) / F.pad on the map (conv.py:279-280, efficientnet.py:108-109)."""
        if left == right == top == bottom == 0:
            return x
        out = self.new(x.N, x.H + top + bottom, x.W + left + right, x.C, dtype=x.dtype)
        out.cmap = x.cmap
        self._use(x, out)
        self.ops.append(lambda plan, ptr, wptr: _lib.call(
            "pcv_zero_pad2d", plan, x.dtype, x.N, x.H, x.W, x.C, ptr(x), x.pitch, left, right, top, bottom, ptr(out),
            out.pitch, None))
        return out

    def dw_pw(self, x: TRef, dwb: nn.Module, pwb: nn.Module, residual: TRef | None, post_act: int | None) -> TRef | None:
        """Depthwise ConvBlock -> pointwise ConvBlock (+ residual, + the unit's activation) as ONE fused kernel
        (include/pcv_b200.h pcv_dw_pw_fused), or None when the pair is outside that kernel's domain."""
        if not _FUSE_DWPW[0] or not _is16(self.dtype) or x.cmap is not None or (residual is not None and residual.cmap is not None):
            return None
        for cb in (dwb, pwb):
            if (type(cb).__name__ != "ConvBlock" or getattr(cb, "use_pad", False) or not cb.normalize
                    or cb.conv.padding_mode != "zeros" or isinstance(cb.conv.padding, str)):
                return None
        cd, cp = dwb.conv, pwb.conv
        if cd.in_channels != x.C or cd.groups != cd.in_channels or cd.out_channels != cd.in_channels:
            return None
        if cp.in_channels != cd.out_channels or cp.groups != 1 or cd.bias is not None or cp.bias is not None:
            return None
        try:
            geo = [(c.kernel_size, _one(c.stride), _one(c.padding), _one(c.dilation)) for c in (cd, cp)]
            act_dw = act_code(dwb.activ) if dwb.activate else ACT_NONE
            act_pw = act_code(pwb.activ) if pwb.activate else ACT_NONE
        except NotImplementedError:
            return None
        if residual is not None:
            if act_pw != ACT_NONE:
                return None                      # the block's own activation would sit between the conv and the add
            act_pw = ACT_NONE if post_act is None else post_act
        elif post_act not in (None, ACT_NONE):
            return None
        (kd, sd, pd, dd), (kp, sp, pp, dp) = geo
        Ho = (x.H + 2 * pd - dd * (kd[0] - 1) - 1) // sd + 1
        Wo = (x.W + 2 * pd - dd * (kd[1] - 1) - 1) // sd + 1
        if Ho <= 0 or Wo <= 0:
            return None
        if residual is not None and (residual.N, residual.H, residual.W, residual.C) != (x.N, Ho, Wo, cp.out_channels):
            return None
        d_dw = ConvDesc(N=x.N, H=x.H, W=x.W, Cin=x.C, Cout=x.C, kh=kd[0], kw=kd[1], stride=sd, pad=pd, dil=dd, groups=x.C,
                        act=act_dw, in_pitch=x.pitch, out_pitch=x.C, res_pitch=0, flags=0)
        d_pw = ConvDesc(N=x.N, H=Ho, W=Wo, Cin=x.C, Cout=cp.out_channels, kh=kp[0], kw=kp[1], stride=sp, pad=pp, dil=dp,
                        groups=1, act=act_pw, in_pitch=x.C, out_pitch=cp.out_channels,
                        res_pitch=residual.pitch if residual is not None else 0, flags=0)
        if not _lib.load().pcv_dw_pw_fusable(C.byref(d_dw), C.byref(d_pw), self.dtype):
            return None
        _check_bn(dwb.bn)
        _check_bn(pwb.bn)
        out = self.new(x.N, Ho, Wo, cp.out_channels)
        offs = []
        for d, cb in ((d_dw, dwb), (d_pw, pwb)):
            wb, bb = C.c_size_t(), C.c_size_t()
            _lib.call("pcv_conv_packed_bytes", C.byref(d), self.dtype, C.byref(wb), C.byref(bb))
            w_off, b_off = self._wblob(wb.value), self._wblob(bb.value)
            self.weight_jobs.append(("conv", (d, cb.conv, cb.bn, 0, w_off, b_off)))
            offs += [w_off, b_off]
        self._use(x, residual, out)
        dtype = self.dtype

        def emit(plan, ptr, wptr):
            _lib.call("pcv_dw_pw_fused", plan, C.byref(d_dw), C.byref(d_pw), dtype, ptr(x), wptr(offs[0]), wptr(offs[1]),
                      wptr(offs[2]), wptr(offs[3]), ptr(residual) if residual is not None else None, ptr(out), None)
        self.ops.append(emit)
        return out

    def exp_dw_pw(self, x: TRef, expb: nn.Module, dwb: nn.Module, pwb: nn.Module, residual: TRef | None,
                  post_act: int | None) -> TRef | None:
        """1x1 expansion ConvBlock -> depthwise ConvBlock -> pointwise ConvBlock (+ residual) as ONE fused kernel
        (include/pcv_b200.h pcv_exp_dw_pw_fused), or None when the triple is outside that kernel's domain."""
        if not _FUSE_XDWPW[0] or not _FUSE_DWPW[0] or not _is16(self.dtype) or x.cmap is not None or (
                residual is not None and residual.cmap is not None):
            return None
        for cb in (expb, dwb, pwb):
            if (type(cb).__name__ != "ConvBlock" or getattr(cb, "use_pad", False) or not cb.normalize
                    or cb.conv.padding_mode != "zeros" or isinstance(cb.conv.padding, str) or cb.conv.bias is not None):
                return None
        ce, cd, cp = expb.conv, dwb.conv, pwb.conv
        if ce.in_channels != x.C or ce.groups != 1 or cd.in_channels != ce.out_channels:
            return None
        if cd.groups != cd.in_channels or cd.out_channels != cd.in_channels or cp.in_channels != cd.out_channels or cp.groups != 1:
            return None
        try:
            geo = [(c.kernel_size, _one(c.stride), _one(c.padding), _one(c.dilation)) for c in (ce, cd, cp)]
            acts = [act_code(cb.activ) if cb.activate else ACT_NONE for cb in (expb, dwb, pwb)]
        except NotImplementedError:
            return None
        act_pw = acts[2]
        if residual is not None:
            if act_pw != ACT_NONE:
                return None
            act_pw = ACT_NONE if post_act is None else post_act
        elif post_act not in (None, ACT_NONE):
            return None
        (ke, se, pe, de), (kd, sd, pd, dd), (kp, sp, pp, dp) = geo
        if sd == 1 and not _XDWPW_S1[0]:
            return None
        Ho = (x.H + 2 * pd - dd * (kd[0] - 1) - 1) // sd + 1
        Wo = (x.W + 2 * pd - dd * (kd[1] - 1) - 1) // sd + 1
        if Ho <= 0 or Wo <= 0:
            return None
        if residual is not None and (residual.N, residual.H, residual.W, residual.C) != (x.N, Ho, Wo, cp.out_channels):
            return None
        cm = ce.out_channels
        d_ex = ConvDesc(N=x.N, H=x.H, W=x.W, Cin=x.C, Cout=cm, kh=ke[0], kw=ke[1], stride=se, pad=pe, dil=de, groups=1,
                        act=acts[0], in_pitch=x.pitch, out_pitch=cm, res_pitch=0, flags=0)
        d_dw = ConvDesc(N=x.N, H=x.H, W=x.W, Cin=cm, Cout=cm, kh=kd[0], kw=kd[1], stride=sd, pad=pd, dil=dd, groups=cm,
                        act=acts[1], in_pitch=cm, out_pitch=cm, res_pitch=0, flags=0)
        d_pw = ConvDesc(N=x.N, H=Ho, W=Wo, Cin=cm, Cout=cp.out_channels, kh=kp[0], kw=kp[1], stride=sp, pad=pp, dil=dp,
                        groups=1, act=act_pw, in_pitch=cm, out_pitch=cp.out_channels,
                        res_pitch=residual.pitch if residual is not None else 0, flags=0)
        if not _lib.load().pcv_exp_dw_pw_fusable(C.byref(d_ex), C.byref(d_dw), C.byref(d_pw), self.dtype):
            return None
        for cb in (expb, dwb, pwb):
            _check_bn(cb.bn)
        out = self.new(x.N, Ho, Wo, cp.out_channels)
        offs = []
        for d, cb in ((d_ex, expb), (d_dw, dwb), (d_pw, pwb)):
            wb, bb = C.c_size_t(), C.c_size_t()
            _lib.call("pcv_conv_packed_bytes", C.byref(d), self.dtype, C.byref(wb), C.byref(bb))
            w_off, b_off = self._wblob(wb.value), self._wblob(bb.value)
            self.weight_jobs.append(("conv", (d, cb.conv, cb.bn, 0, w_off, b_off)))
            offs += [w_off, b_off]
        self._use(x, residual, out)
        dtype = self.dtype

        def emit(plan, ptr, wptr):
            _lib.call("pcv_exp_dw_pw_fused", plan, C.byref(d_ex), C.byref(d_dw), C.byref(d_pw), dtype, ptr(x),
                      *(wptr(o) for o in offs), ptr(residual) if residual is not None else None, ptr(out), None)
        self.ops.append(emit)
        return out

    def linear(self, x: TRef, fc: nn.Linear, out_f32: bool = True) -> TRef:
        """nn.Linear on pooled features == 1x1 conv on a 1x1 map (resnet.py:320-322,335-336)."""
        if (x.H, x.W) != (1, 1):
            raise ValueError("Linear expects a 1x1 feature map (the reference flattens [N,C,1,1])")
        shim = _ConvShim(fc.weight.view(fc.out_features, fc.in_features, 1, 1), fc.bias)
        return self.conv(x, shim, None, ACT_NONE, out_f32=out_f32)

    def maxpool(self, x: TRef, k: int, stride: int, pad: int) -> TRef:
        Ho = (x.H + 2 * pad - k) // stride + 1
        Wo = (x.W + 2 * pad - k) // stride + 1
        st = self.stem
        if (self.allow_stem_pool and st is not None and x is st.get("out") and st.get("pool_ok")
                and st["idx"] == len(self.ops) - 1 and x.buf.last == st["idx"] and (k, stride, pad) == (3, 2, 1)):
            # ResInitBlock (resnet.py:255-263): conv7x7_block -> MaxPool2d(3, 2, 1).  The stem kernel reduces the pooled
            # rows from its staged conv tile, so the conv map is never written (include/pcv_b200.h PCV_CONV_POOL3S2).
            pooled = self.new(x.N, Ho, Wo, x.C, dtype=x.dtype)
            pooled.buf.first = pooled.buf.last = st["idx"]
            st["desc"].flags |= _lib.CONV_POOL3S2
            st["desc"].out_pitch = pooled.pitch
            st["target"][0] = pooled
            st["fused_conv_buf"] = x.buf
            x.buf.nbytes = 0
            return pooled
        out = self.new(x.N, Ho, Wo, x.C, dtype=x.dtype)
        out.cmap = x.cmap
        self._use(x, out)
        self.ops.append(lambda plan, ptr, wptr: _lib.call(
            "pcv_maxpool2d", plan, x.dtype, x.N, x.H, x.W, x.C, k, stride, pad, ptr(x), x.pitch, ptr(out), out.pitch,
            None))
        return out

    def gap(self, x: TRef, out_dtype: int | None = None) -> TRef:
        out_dtype = x.dtype if out_dtype is None else out_dtype
        out = self.new(x.N, 1, 1, x.C, dtype=out_dtype)
        out.cmap = x.cmap
        self._use(x, out)
        self.ops.append(lambda plan, ptr, wptr: _lib.call(
            "pcv_global_avgpool", plan, x.dtype, x.N, x.H * x.W, x.C, ptr(x), x.pitch, ptr(out), out_dtype, None))
        return out

    def adaptive_pool(self, x: TRef, out_h: int, out_w: int) -> TRef:
        """nn.AdaptiveAvgPool2d((out_h, out_w)) (pspnet.py:71); (1, 1) is the global pool."""
        if (out_h, out_w) == (1, 1):
            return self.gap(x)
        out = self.new(x.N, out_h, out_w, x.C, dtype=x.dtype)
        out.cmap = x.cmap
        self._use(x, out)
        self.ops.append(lambda plan, ptr, wptr: _lib.call(
            "pcv_adaptive_avgpool", plan, x.dtype, x.N, x.H, x.W, x.C, ptr(x), x.pitch, out_h, out_w, ptr(out), None))
        return out

    def se_gate(self, pooled: TRef, w1: torch.Tensor, b1, w2: torch.Tensor, b2, mid_act: int, out_act: int) -> TRef:
        """SEBlock excite (att.py:99-102): gate = out_act(W2 @ mid_act(W1 @ pooled + b1) + b2), all fp32.  `pooled` may be
        narrower than the gate (w1 [mid, pooled.C], w2 [C, mid]) when the unit's last 1x1 conv was folded into W1."""
        cmid = w1.shape[0]
        if pooled.cmap is not None:
            # virtual channel padding: W1's columns / W2's rows (and b2) move to the storage positions, zeros elsewhere - the
            # padding channels' gates multiply exact zeros
            idx = torch.tensor(pooled.cmap, dtype=torch.long, device=w1.device)
            if w1.shape[1] != len(pooled.cmap) or w2.shape[0] != len(pooled.cmap):
                raise NotImplementedError("SE block on a padded tensor must gate the tensor it pools")
            w1s = torch.zeros(cmid, pooled.C, dtype=w1.dtype, device=w1.device)
            w1s[:, idx] = w1.detach()
            w2s = torch.zeros(pooled.C, cmid, dtype=w2.dtype, device=w2.device)
            w2s[idx] = w2.detach()
            if b2 is not None:
                b2s = torch.zeros(pooled.C, dtype=b2.dtype, device=b2.device)
                b2s[idx] = b2.detach()
                b2 = b2s
            w1, w2 = w1s, w2s
        N, Cin, Cc = pooled.N, pooled.C, w2.shape[0]
        gate = self.new(N, 1, 1, Cc, dtype=F32, extra_bytes=N * cmid * 4)
        gate.cmap = pooled.cmap
        offs = []
        for t in (w1, b1, w2, b2):
            offs.append(None if t is None else self._wblob(t.numel() * 4))
            if t is not None:
                self.weight_jobs.append(("raw", (t, offs[-1])))
        self._use(pooled, gate)

        def emit(plan, ptr, wptr):
            _lib.call("pcv_se_excite_ex", plan, N, Cin, cmid, Cc, ptr(pooled), wptr(offs[0]),
                      wptr(offs[1]) if offs[1] is not None else None, wptr(offs[2]),
                      wptr(offs[3]) if offs[3] is not None else None, mid_act, out_act, ptr(gate), None)
        self.ops.append(emit)
        return gate

    def se_scale(self, x: TRef, gate: TRef, identity: TRef | None, act: int) -> TRef:
        for t in (x, identity):
            if t is not None and t.pitch != t.C:
                raise NotImplementedError("SE scale needs dense tensors")
        if identity is not None and identity.cmap != x.cmap:
            raise ValueError("SE identity has a different channel layout")
        out = self.new(x.N, x.H, x.W, x.C, dtype=x.dtype)
        out.cmap = x.cmap
        self._use(x, gate, identity, out)
        self.ops.append(lambda plan, ptr, wptr: _lib.call(
            "pcv_se_scale_add_act", plan, x.dtype, x.N, x.H * x.W, x.C, ptr(x), ptr(gate),
            ptr(identity) if identity is not None else None, act, ptr(out), None))
        return out

    def add_act(self, a: TRef, b: TRef, act: int) -> TRef:
        if (a.N, a.H, a.W, a.C) != (b.N, b.H, b.W, b.C) or a.pitch != a.C or b.pitch != b.C:
            raise ValueError("add needs two dense tensors of the same shape")
        if a.cmap != b.cmap:
            raise ValueError("add of two tensors with different channel layouts")
        out = self.new(a.N, a.H, a.W, a.C, dtype=a.dtype)
        out.cmap = a.cmap
        self._use(a, b, out)
        self.ops.append(lambda plan, ptr, wptr: _lib.call(
            "pcv_add_act", plan, a.dtype, a.N * a.H * a.W * a.C, ptr(a), ptr(b), act, ptr(out), None))
        return out

    def affine_act(self, x: TRef, scale: torch.Tensor | None, shift: torch.Tensor | None, act: int = ACT_NONE,
                   slope: torch.Tensor | None = None) -> TRef:
        """y = act(x * scale[c] + shift[c]), negative side times slope[c] (include/pcv_b200.h pcv_channel_affine_act): the
        BN -> ReLU pre-activation of PreConvBlock / PreResActivation, nn.PReLU, LeakyReLU behind a depthwise conv."""
        if x.C % 8 != 0:
            raise NotImplementedError(f"a stand-alone normalisation / activation pass needs C % 8 == 0, got {x.C}")
        out = self.new(x.N, x.H, x.W, x.C, dtype=x.dtype)
        out.cmap = x.cmap
        if x.cmap is not None and act in (ACT_SIGMOID, ACT_HSIGMOID):
            raise NotImplementedError("act(0) != 0 on a tensor with virtual channel padding")
        offs = []
        for k, t in enumerate((scale, shift, slope)):
            if t is None:
                offs.append(None)
                continue
            if t.numel() != x.creal:
                raise ValueError(f"per-channel vector of {t.numel()} values for {x.creal} channels")
            if x.cmap is not None:   # scatter to the storage positions: scale 1 / shift 0 / slope 1 on the padding (zeros stay zeros)
                full = torch.full((x.C,), 0.0 if k == 1 else 1.0, dtype=torch.float32, device=t.device)
                full[torch.tensor(x.cmap, dtype=torch.long, device=t.device)] = t.detach().float()
                t = full
            offs.append(self._wblob(x.C * 4))
            self.weight_jobs.append(("raw", (t, offs[-1])))
        self._use(x, out)

        def emit(plan, ptr, wptr):
            _lib.call("pcv_channel_affine_act", plan, x.dtype, x.N * x.H * x.W, x.C, ptr(x), x.pitch,
                      *(wptr(o) if o is not None else None for o in offs[:3]), act, ptr(out), out.pitch, None)
        self.ops.append(emit)
        return out

    def activation(self, x: TRef, m: nn.Module) -> TRef:
        """A stand-alone activation module on a map (PReLU / LeakyReLU / any pcv_act code) as one pass."""
        if isinstance(m, nn.PReLU):
            w = m.weight.detach().float()
            return self.affine_act(x, None, None, ACT_NONE, (w.expand(x.creal) if w.numel() == 1 else w).contiguous())
        if isinstance(m, nn.LeakyReLU):
            return self.affine_act(x, None, None, ACT_NONE,
                                   torch.full((x.creal,), float(m.negative_slope), dtype=torch.float32, device=self.device))
        return self.affine_act(x, None, None, act_code(m))

    def _edge_out(self, x: TRef, H: int, W: int, name: str, launch) -> TRef:
        """An fp32 NCHW tensor the reference returns (freshly allocated, caller-owned): produced per call by
        `launch(plan=None, ptr, out_ptr, stream)` after the plan, straight into a new torch tensor - no arena storage."""
        self._use(x)
        x.buf.pinned = True   # read after the whole plan has run
        out = TRef(x.N, H, W, x.C, x.C, F32, Buf(nbytes=0, first=len(self.ops)), 0, "nchw")
        out.tail = {"launch": launch, "name": name,
                    "bytes": float(x.N * x.C * (_esize(x.dtype) * x.H * x.W + 4 * H * W))}
        return out

    def bilinear(self, x: TRef, Hout: int, Wout: int, out: TRef | None = None, nchw_f32: bool = False) -> TRef:
        if x.cmap is not None:
            x = self.compact(x)
        if nchw_f32:
            return self._edge_out(x, Hout, Wout, f"bilinear C={x.C} {x.H}x{x.W}->{Hout}x{Wout} nchw_f32 (edge)",
                                  lambda ptr, optr, stream: _lib.call(
                "pcv_bilinear_upsample_ac", None, x.dtype, x.N, x.H, x.W, x.C, ptr(x), x.pitch, Hout, Wout, optr,
                x.C, 1, stream))
        if out is None:
            out = self.new(x.N, Hout, Wout, x.C, dtype=x.dtype)
        self._use(x, out)
        self.ops.append(lambda plan, ptr, wptr: _lib.call(
            "pcv_bilinear_upsample_ac", plan, x.dtype, x.N, x.H, x.W, x.C, ptr(x), x.pitch, Hout, Wout, ptr(out),
            out.pitch, 1 if nchw_f32 else 0, None))
        return out

    def relayout(self, x: TRef, cmap: list | None, storage: int | None = None) -> TRef:
        """Move the real channels of `x` to the storage positions `cmap` (None: dense) of a `storage`-channel tensor: a 1x1
        convolution whose weight is the identity on the real channels (exact in every tier: products by 1.0, sums with
        zeros).  The one extra pass a torch.split / torch.cat costs when producer and consumer disagree on the padding."""
        n = x.creal
        want = list(cmap) if cmap is not None else list(range(n))
        have = list(x.cmap) if x.cmap is not None else list(range(n))
        storage = _rup(max(want) + 1, 8) if storage is None else storage
        if want == have and storage == x.C:
            return x
        shim = _ConvShim(torch.eye(n, dtype=torch.float32, device=self.device).view(n, n, 1, 1), None)
        return self.conv(x, shim, None, ACT_NONE, out=self.new(x.N, x.H, x.W, storage), out_cmap=want)

    def compact(self, x: TRef) -> TRef:
        """Gather the real channels of a tensor with virtual channel padding into a dense tensor."""
        return x if x.cmap is None else self.relayout(x, None)

    def egress(self, x: TRef) -> TRef:
        if x.cmap is not None and list(x.cmap) != list(range(x.creal)):
            x = self.compact(x)
        creal = x.creal   # a prefix map (padding at the end only) is just a narrower tensor at the same pitch
        out = self._edge_out(x, x.H, x.W, f"nhwc_to_nchw_f32 C={creal} {x.H}x{x.W} (edge)", lambda ptr, optr, stream: _lib.call(
            "pcv_nhwc_to_nchw_f32", None, x.dtype, x.N, creal, x.H, x.W, ptr(x), x.pitch, optr, stream))
        out.C = out.pitch = creal
        return out


class _ConvShim:
    """Presents an nn.Linear (or SE fc) as the nn.Conv2d attribute set Builder.conv reads."""

    def __init__(self, weight4d: torch.Tensor, bias):
        self.weight, self.bias = weight4d, bias
        self.out_channels, self.in_channels = weight4d.shape[0], weight4d.shape[1]
        self.kernel_size, self.stride, self.padding, self.dilation = (1, 1), (1, 1), (0, 0), (1, 1)
        self.groups, self.padding_mode = 1, "zeros"


# ---------------------------------------------------------------------------------------------------------------
# lowering rules, keyed by the reference's class names
# ---------------------------------------------------------------------------------------------------------------
LOWER: dict[str, Callable] = {}


def lowers(*names):
    def deco(fn):
        for n in names:
            LOWER[n] = fn
        return fn
    return deco


def lower(b: Builder, m: nn.Module, x, **kw):
    fn = LOWER.get(type(m).__name__)
    if fn is None and type(m) is nn.Sequential:
        fn = _lower_sequential
    if fn is None:
        raise NotImplementedError(
            f"module {type(m).__name__} is outside the B200 eval path (no CPU fallback; see SURVEY.md section 8)")
    return fn(b, m, x, **kw)


def _lower_sequential(b, m, x, **kw):
    mods = list(m.children())
    for i, child in enumerate(mods):
        x = lower(b, child, x, **(kw if i == len(mods) - 1 else {}))
    return x


@lowers("Dropout", "Identity")
def _lower_identity(b, m, x, **kw):
    if isinstance(m, nn.Dropout) and m.training:
        raise RuntimeError("pytorchcv_b200 is an eval-mode path: call net.eval() first")
    return x


@lowers("Conv2d")
def _lower_conv2d(b, m, x, out=None, out_f32=False, **kw):
    """bare conv1x1 / conv3x3 (conv.py:89-164): no BN, no activation."""
    return b.conv(x, m, None, ACT_NONE, out=out, out_f32=out_f32)


@lowers("Linear")
def _lower_linear(b, m, x, **kw):
    return b.linear(x, m)


@lowers("ConvBlock")
def _lower_convblock(b, m, x, residual=None, post_act=None, out=None, pad_lrtb=None, out_cmap=None, **kw):
    """ConvBlock.forward (conv.py:278-286); `residual`/`post_act` carry the enclosing unit's add + activation.  A block built
    with a 4-tuple padding applies nn.ZeroPad2d first (conv.py:245-249,279-280); `pad_lrtb` is the caller's F.pad (tf_mode)."""
    if getattr(m, "use_pad", False):
        if pad_lrtb is not None:
            raise NotImplementedError("ZeroPad2d ConvBlock behind an explicit F.pad")
        pad_lrtb = tuple(int(v) for v in m.pad.padding)
    bn = m.bn if m.normalize else None
    post = ACT_NONE if post_act is None else post_act
    if m.activate and _standalone_act(m.activ, m.conv):
        # nn.PReLU (per-channel slopes) / LeakyReLU behind a depthwise conv: one pass behind the conv
        if out is not None:
            raise NotImplementedError("a ConvBlock with a stand-alone activation cannot write into a concat slice")
        y = b.activation(b.conv(x, m.conv, bn, ACT_NONE, pad_lrtb=pad_lrtb, out_cmap=out_cmap), m.activ)
        return b.add_act(y, residual, post) if residual is not None else y
    act = act_code(m.activ) if m.activate else ACT_NONE
    act_a = act_param(m.activ) if m.activate else 0.0
    if residual is None and post_act is None:
        return b.conv(x, m.conv, bn, act, out=out, pad_lrtb=pad_lrtb, act_a=act_a, out_cmap=out_cmap)
    if act == ACT_NONE:
        return b.conv(x, m.conv, bn, post, residual=residual, out=out, pad_lrtb=pad_lrtb, out_cmap=out_cmap)   # fused: act(conv + residual)
    y = b.conv(x, m.conv, bn, act, pad_lrtb=pad_lrtb, act_a=act_a, out_cmap=out_cmap)   # block has its own activation
    return b.add_act(y, residual, post) if residual is not None else y


@lowers("DwsConvBlock")
def _lower_dws(b, m, x, **kw):
    """DwsConvBlock.forward (conv.py:605-608): depthwise ConvBlock then pointwise ConvBlock."""
    return _dw_then_pw(b, m.dw_conv, m.pw_conv, x, **kw)


@lowers("MaxPool2d")
def _lower_maxpool(b, m, x, **kw):
    if m.ceil_mode or _one(m.dilation) != 1:
        raise NotImplementedError("MaxPool2d with ceil_mode / dilation is outside the B200 eval path")
    return b.maxpool(x, _one(m.kernel_size), _one(m.stride if m.stride is not None else m.kernel_size),
                     _one(m.padding))


@lowers("AvgPool2d")
def _lower_avgpool(b, m, x, **kw):
    """The reference only uses AvgPool2d(7, stride=1) on a 7x7 map, i.e. a global mean (resnet.py:316-318)."""
    k = _one(m.kernel_size)
    if k != x.H or k != x.W or _one(m.padding) != 0:
        raise NotImplementedError(f"AvgPool2d({k}) on a {x.H}x{x.W} map is not a global pool (SURVEY appendix B)")
    return b.gap(x)


@lowers("AdaptiveAvgPool2d")
def _lower_adaptive(b, m, x, **kw):
    size = m.output_size if isinstance(m.output_size, (tuple, list)) else (m.output_size, m.output_size)
    if None in size:
        raise NotImplementedError("AdaptiveAvgPool2d with a None output size is outside the B200 eval path")
    return b.adaptive_pool(x, int(size[0]), int(size[1]))


def _se_parts(b, m):
    if m.use_conv:
        c1, c2 = m.conv1, m.conv2
        w1, w2 = c1.weight.view(c1.out_channels, -1), c2.weight.view(c2.out_channels, -1)
    else:
        c1, c2 = m.fc1, m.fc2
        w1, w2 = c1.weight, c2.weight
    return w1, c1.bias, w2, c2.bias, act_code(m.activ), act_code(m.sigmoid)


@lowers("SEBlock")
def _lower_se(b, m, x, identity=None, post_act=ACT_NONE, **kw):
    """SEBlock.forward (att.py:94-105): x * sigmoid(W2 relu(W1 mean(x) + b1) + b2) [+ identity, act]."""
    w1, b1, w2, b2, mid_act, out_act = _se_parts(b, m)
    pooled = b.gap(x, out_dtype=F32)
    gate = b.se_gate(pooled, w1, b1, w2, b2, mid_act, out_act)
    return b.se_scale(x, gate, identity, post_act)


@lowers("ResBlock", "ResBottleneck", "ResNeXtBottleneck", "SENetBottleneck")
def _lower_resbody(b, m, x, residual=None, post_act=None, **kw):
    """conv1 -> conv2 [-> conv3] (resnet.py:63-66,136-140; resnext.py:56-59); the last conv takes the fusion."""
    convs = [m.conv1, m.conv2] + ([m.conv3] if hasattr(m, "conv3") else [])
    if len(convs) == 3 and residual is not None and post_act is not None:
        y1 = lower(b, convs[0], x)
        fused = b.bottleneck_tail(y1, convs[1], convs[2], residual, post_act)   # 3x3 -> 1x1 + add + act in one kernel
        if fused is not None:
            return fused
        return lower(b, convs[2], lower(b, convs[1], y1), residual=residual, post_act=post_act)
    for c in convs[:-1]:
        x = lower(b, c, x)
    return lower(b, convs[-1], x, residual=residual, post_act=post_act)


# A unit's projection shortcut folded into its last 1x1 conv (pcv_conv1x1_dual); PCV_DUAL_IDENTITY=0 keeps the separate
# identity_conv kernel and the residual add in conv3's epilogue
_DUAL_IDENTITY = [os.environ.get("PCV_DUAL_IDENTITY", "1") != "0"]


def set_dual_identity(enabled: bool) -> None:
    _DUAL_IDENTITY[0] = bool(enabled)


# the gated variant (SE units) runs 128-wide tiles with two accumulators: it pays on the bandwidth-bound early stages only
# (SE-ResNeXt-50 bs256, identity conv + gated conv3 -> one op: @56x56 334 -> 234 us, @28x28 200 -> 193, @14x14 136 -> 140,
# @7x7 123 -> 127), so the default takes it for output maps of >= 28 x 28
_DUAL_GATE_MIN_HW = [int(os.environ.get("PCV_DUAL_GATE_MIN_HW", "784"))]


def set_dual_gate_min_hw(pixels: int) -> None:
    _DUAL_GATE_MIN_HW[0] = int(pixels)


@lowers("ResUnit", "ResNeXtUnit")
def _lower_resunit(b, m, x, **kw):
    """ResUnit.forward (resnet.py:221-229): act(body(x) + (identity_conv(x) | x)).

    With a projection shortcut behind a bottleneck body, conv3 and identity_conv are both linear 1x1 ConvBlocks landing on the
    same grid: act(W3 y2 + b3 + Wid x[::s] + bid) is one GEMM over [y2 ; x[::s]] (Builder.conv_dual)."""
    body = m.body
    if (m.resize_identity and _DUAL_IDENTITY[0] and _is16(b.dtype) and hasattr(body, "conv3")
            and type(body).__name__ in ("ResBottleneck", "ResNeXtBottleneck") and not _FUSE_TAIL[0]):
        y2 = lower(b, body.conv2, lower(b, body.conv1, x))
        fused = b.conv_dual(y2, body.conv3, x, m.identity_conv, act_code(m.activ))
        if fused is not None:
            return fused
        return lower(b, body.conv3, y2, residual=lower(b, m.identity_conv, x), post_act=act_code(m.activ))
    identity = lower(b, m.identity_conv, x) if m.resize_identity else x
    return lower(b, body, x, residual=identity, post_act=act_code(m.activ))


# SE scale + identity + activation in the epilogue of the unit's last 1x1 conv (PCV_CONV_SE_GATE); PCV_SE_GATE_FUSE=0 keeps the
# separate pcv_se_scale_add_act pass
_SE_GATE_FUSE = [os.environ.get("PCV_SE_GATE_FUSE", "1") != "0"]


def set_se_gate_fuse(enabled: bool) -> None:
    _SE_GATE_FUSE[0] = bool(enabled)


def _fold_conv3_into_se(conv3, se):
    """mean_HW(conv3(y)) == W3' mean_HW(y) + b3' for a 1x1 stride-1 ConvBlock without activation (BN folded into W3', b3'):
    returns (W1 W3', W1 b3' + b1) so that the SE squeeze can pool conv3's INPUT (fewer channels), or None."""
    c = getattr(conv3, "conv", None)
    if (type(conv3).__name__ != "ConvBlock" or conv3.activate or getattr(conv3, "use_pad", False) or c is None
            or c.kernel_size != (1, 1) or _one(c.stride) != 1 or _one(c.padding) != 0 or c.groups != 1
            or not _SE_FOLD or c.in_channels >= c.out_channels):
        return None
    with torch.no_grad():
        w3 = c.weight.detach().double().view(c.out_channels, c.in_channels)
        b3 = c.bias.detach().double() if c.bias is not None else torch.zeros(c.out_channels, dtype=torch.float64,
                                                                              device=w3.device)
        if conv3.normalize:
            bn = conv3.bn
            _check_bn(bn)
            g = bn.weight.detach().double() if bn.weight is not None else torch.ones_like(bn.running_var.double())
            be = bn.bias.detach().double() if bn.bias is not None else torch.zeros_like(g)
            sc = g / torch.sqrt(bn.running_var.detach().double() + bn.eps)
            w3, b3 = w3 * sc[:, None], (b3 - bn.running_mean.detach().double()) * sc + be
        if se.use_conv:
            w1 = se.conv1.weight.detach().double().view(se.conv1.out_channels, -1)
            b1 = se.conv1.bias
        else:
            w1, b1 = se.fc1.weight.detach().double(), se.fc1.bias
        b1 = b1.detach().double() if b1 is not None else torch.zeros(w1.shape[0], dtype=torch.float64, device=w1.device)
        return (w1 @ w3).float().contiguous(), (w1 @ b3 + b1).float().contiguous()


@lowers("SEResNeXtUnit", "SEResUnit", "SENetUnit")
def _lower_seresnext_unit(b, m, x, **kw):
    """SEResNeXtUnit.forward (seresnext.py:57-66): relu(se(body(x)) + identity).

    When the body ends in a linear 1x1 ConvBlock (ResNeXtBottleneck / ResBottleneck / SENetBottleneck conv3) the SE squeeze
    is taken on that conv's INPUT: global mean commutes with the 1x1 conv + BN, whose weights fold into the first SE FC on
    the host.  The squeeze pass then reads the bottleneck width instead of the unit width (half the bytes in SE-ResNeXt,
    a quarter in SE-ResNet) and sees unrounded values of conv3's output."""
    body, se = m.body, m.se
    folded = _fold_conv3_into_se(body.conv3, se) if hasattr(body, "conv3") else None
    if folded is None:
        identity = lower(b, m.identity_conv, x) if m.resize_identity else x
        y = lower(b, body, x)
        return lower(b, se, y, identity=identity, post_act=act_code(m.activ))
    # a projection shortcut that can ride on conv3 as the second half of its K dimension (Builder.conv_dual) is never computed
    # on its own; otherwise it keeps its place in front of the body
    dual = m.resize_identity and _SE_GATE_FUSE[0] and _DUAL_IDENTITY[0] and _is16(b.dtype)
    identity = None if dual else (lower(b, m.identity_conv, x) if m.resize_identity else x)
    y2 = lower(b, body.conv2, lower(b, body.conv1, x))
    pooled = b.gap(y2, out_dtype=F32)
    _, _, w2, b2, mid_act, out_act = _se_parts(b, se)
    gate = b.se_gate(pooled, folded[0], folded[1], w2, b2, mid_act, out_act)
    # the gate exists BEFORE conv3 runs (it was squeezed from conv3's input), so the SE scale, the identity add and the unit's
    # activation ride on conv3's epilogue: conv3's output and the scale pass's read of it never touch HBM
    if dual:
        fused = b.conv_dual(y2, body.conv3, x, m.identity_conv, act_code(m.activ), gate=gate)
        if fused is not None:
            return fused
        identity = lower(b, m.identity_conv, x)
    if _SE_GATE_FUSE[0]:
        c3 = body.conv3
        fused = b.conv(y2, c3.conv, c3.bn if c3.normalize else None, act_code(m.activ), residual=identity, gate=gate)
        if fused is not None:
            return fused
    y3 = lower(b, body.conv3, y2)
    return b.se_scale(y3, gate, identity, act_code(m.activ))


@lowers("ResInitBlock")
def _lower_resinit(b, m, x, **kw):
    return lower(b, m.pool, lower(b, m.conv, x))


@lowers("SEInitBlock")
def _lower_seinit(b, m, x, **kw):
    for c in (m.conv1, m.conv2, m.conv3):
        x = lower(b, c, x)
    return lower(b, m.pool, x)


@lowers("LinearBottleneck")
def _lower_linear_bottleneck(b, m, x, **kw):
    """LinearBottleneck.forward (mobilenetv2.py:62-71): [1x1 expand] -> dw3x3 -> 1x1 linear (+x), no final act."""
    if m.use_exp_conv:
        return _exp_dw_pw(b, m.conv1, m.conv2, m.conv3, x, residual=x if m.residual else None, post_act=None)
    return _dw_then_pw(b, m.conv2, m.conv3, x, residual=x if m.residual else None, post_act=None)


# ---- pre-activation family (conv.py:652-732, preresnet.py) ------------------------------------------------------------
def _pre_act(b, m, x):
    """The BN -> activation half of a PreConvBlock (conv.py:717-721) as one stand-alone pass over `x`."""
    scale = shift = None
    if m.normalize:
        scale, shift = _bn_scale_shift(m.bn)
    if m.activate and _standalone_act(m.activ):
        return b.activation(b.affine_act(x, scale, shift) if scale is not None else x, m.activ)
    if m.activate and isinstance(m.activ, nn.LeakyReLU):
        slope = torch.full((x.C,), float(m.activ.negative_slope), dtype=torch.float32)
        return b.affine_act(x, scale, shift, ACT_NONE, slope)
    act = act_code(m.activ) if m.activate else ACT_NONE
    if scale is None and act == ACT_NONE:
        return x
    return b.affine_act(x, scale, shift, act)


@lowers("PreConvBlock")
def _lower_preconv(b, m, x, **kw):
    """PreConvBlock.forward (conv.py:717-731): BN -> ReLU -> conv; with return_preact the activated map is the second result
    (PreResUnit feeds it to its identity convolution, preresnet.py:157-163)."""
    pre = _pre_act(b, m, x)
    y = b.conv(pre, m.conv, None, ACT_NONE)
    return (y, pre) if m.return_preact else y


def _pre_chain(b, blocks, x, residual=None):
    """PreConvBlocks in series: block i+1's BN -> activation is a per-output-channel affine + activation of block i's conv, so
    it rides on that conv's epilogue (folded weights / bias); only the first block's pre-activation needs its own pass.
    `residual`: a tensor, or a function of the first block's pre-activated input (PreResUnit's projection shortcut).
    Returns (result of the last conv [+ residual], first block's pre-activated input)."""
    pre = _pre_act(b, blocks[0], x)
    if callable(residual):
        residual = residual(pre)
    y = pre
    for i, blk in enumerate(blocks):
        nxt = blocks[i + 1] if i + 1 < len(blocks) else None
        if nxt is None:
            y = b.conv(y, blk.conv, None, ACT_NONE, residual=residual)
        elif _standalone_act(nxt.activ if nxt.activate else None):
            y = _pre_act(b, nxt, b.conv(y, blk.conv, None, ACT_NONE))
        else:
            y = b.conv(y, blk.conv, nxt.bn if nxt.normalize else None, act_code(nxt.activ) if nxt.activate else ACT_NONE,
                       act_a=act_param(nxt.activ) if nxt.activate else 0.0)
    return y, pre


@lowers("PreResBlock", "PreResBottleneck")
def _lower_preres_body(b, m, x, residual=None, **kw):
    """PreResBlock.forward / PreResBottleneck.forward (preresnet.py:56-59, 98-102): conv1 (returns its pre-activation) ->
    conv2 [-> conv3]; result (x, x_pre_activ) like the reference."""
    blocks = [m.conv1, m.conv2] + ([m.conv3] if hasattr(m, "conv3") else [])
    return _pre_chain(b, blocks, x, residual=residual)


@lowers("PreResUnit")
def _lower_preres_unit(b, m, x, **kw):
    """PreResUnit.forward (preresnet.py:157-163): body(x) + (identity_conv(pre-activated x) | x), no activation after the add.
    The add rides on the last conv's epilogue; the projection reads the pre-activated map the body's first block produced."""
    shortcut = (lambda pre: lower(b, m.identity_conv, pre)) if m.resize_identity else x
    return lower(b, m.body, x, residual=shortcut)[0]


@lowers("PreResInitBlock")
def _lower_preres_init(b, m, x, **kw):
    """PreResInitBlock.forward (preresnet.py:195-200): bare conv7x7 -> BN -> ReLU -> max pool == ResInitBlock's arithmetic."""
    return lower(b, m.pool, b.conv(x, m.conv, m.bn, act_code(m.activ)))


@lowers("PreResActivation")
def _lower_preres_activation(b, m, x, **kw):
    """PreResActivation.forward (preresnet.py:218-221): the network's last BN -> ReLU."""
    scale, shift = _bn_scale_shift(m.bn)
    return b.affine_act(x, scale, shift, act_code(m.activ))


# ---- GhostNet (ghostnet.py): torch.cat of a 1x1 conv and a cheap depthwise conv of it ---------------------------------------
@lowers("GhostConvBlock")
def _lower_ghost_conv(b, m, x, out_cmap=None, **kw):
    """GhostConvBlock.forward (ghostnet.py:57-60): x = main_conv(x); cat(x, cheap_conv(x)).  Both halves are written straight
    into channel slices of one buffer; a half whose width is not a multiple of 8 is padded to 8 with zero channels (TRef.cmap)
    so that the cheap depthwise conv and every consumer see 16-byte aligned slices."""
    if kw.get("residual") is not None or kw.get("post_act") is not None or out_cmap is not None:
        raise NotImplementedError("GhostConvBlock takes no fused residual")
    main_c, cheap_c = m.main_conv.conv.out_channels, m.cheap_conv.conv.out_channels
    if main_c != cheap_c:
        raise NotImplementedError("GhostConvBlock with an odd width")
    ms = _rup(main_c, 8)
    conv0 = m.main_conv.conv
    Ho = (x.H + 2 * _one(conv0.padding) - _one(conv0.dilation) * (conv0.kernel_size[0] - 1) - 1) // _one(conv0.stride) + 1
    Wo = (x.W + 2 * _one(conv0.padding) - _one(conv0.dilation) * (conv0.kernel_size[1] - 1) - 1) // _one(conv0.stride) + 1
    cat = b.new(x.N, Ho, Wo, 2 * ms)
    xm = lower(b, m.main_conv, x, out=Builder.view(cat, 0, ms))
    y = lower(b, m.cheap_conv, xm, out=Builder.view(cat, ms, ms))
    if y.buf is not cat.buf or xm.buf is not cat.buf:
        raise NotImplementedError("GhostConvBlock halves must write into the concat buffer")
    cat.cmap = None if ms == main_c else list(range(main_c)) + [ms + i for i in range(cheap_c)]
    return cat


@lowers("GhostExpBlock")
def _lower_ghost_exp(b, m, x, **kw):
    """GhostExpBlock.forward (ghostnet.py:114-121): exp_conv -> [depthwise stride-2 conv] -> [SE] -> pw_conv."""
    y = lower(b, m.exp_conv, x)
    if m.use_dw_conv:
        y = lower(b, m.dw_conv, y)
    if m.use_se:
        y = lower(b, m.se, y)
    return lower(b, m.pw_conv, y)


@lowers("GhostUnit")
def _lower_ghost_unit(b, m, x, **kw):
    """GhostUnit.forward (ghostnet.py:167-174): body(x) + (identity_conv(x) | x), nothing after the add.  The projection
    shortcut's pointwise conv writes its channels in the body's (padded) concat layout so that the add is elementwise."""
    y = lower(b, m.body, x)
    if m.resize_identity:
        idc = m.identity_conv
        identity = lower(b, idc.pw_conv, lower(b, idc.dw_conv, x), out=b.new(y.N, y.H, y.W, y.C),
                         out_cmap=y.cmap if y.cmap is not None else list(range(y.C)))
    else:
        identity = b.relayout(x, y.cmap, y.C)   # a no-op inside the network: the previous unit left the same layout
    return b.add_act(y, identity, ACT_NONE)


@lowers("GhostClassifier")
def _lower_ghost_classifier(b, m, x, **kw):
    """GhostClassifier.forward (ghostnet.py:203-206): 1x1 ConvBlock then a bare 1x1 conv with bias on the 1x1 map (fp32 logits)."""
    return b.conv(lower(b, m.conv1, x), m.conv2, None, ACT_NONE, out_f32=True)


@lowers("GhostNet")
def _lower_ghostnet(b, m, x, **kw):
    """GhostNet.forward (ghostnet.py:298-302): features -> classifier convs -> view."""
    return _flat(lower(b, m.output, lower(b, m.features, x)))


# ---- MixNet (mixnet.py): torch.split -> one conv per part (mixed kernel sizes) -> torch.cat -> ONE BatchNorm -----------------
def _seg_layout(sizes):
    """Channel layout of a concat of parts: part i starts at the sum of the 8-rounded widths before it.  Returns (cmap | None,
    storage width, part offsets)."""
    offs, off = [], 0
    for n in sizes:
        offs.append(off)
        off += _rup(n, 8)
    cmap = [o + i for o, n in zip(offs, sizes) for i in range(n)]
    return (None if cmap == list(range(sum(sizes))) and off == sum(sizes) else cmap), off, offs


def _bn_slice(bn: nn.BatchNorm2d, lo: int, hi: int) -> nn.BatchNorm2d:
    """Channels [lo, hi) of an eval BatchNorm2d as a BatchNorm2d of their own (MixConvBlock normalises the concat, mixnet.py:151-157;
    a changed parameter recompiles the plan, so a copy is as good as a view)."""
    _check_bn(bn)
    part = nn.BatchNorm2d(hi - lo, eps=bn.eps, affine=bn.weight is not None).to(bn.running_mean.device).eval()
    with torch.no_grad():
        part.running_mean.copy_(bn.running_mean[lo:hi])
        part.running_var.copy_(bn.running_var[lo:hi])
        if bn.weight is not None:
            part.weight.copy_(bn.weight[lo:hi])
            part.bias.copy_(bn.bias[lo:hi])
    return part


@lowers("MixConvBlock")
def _lower_mixconv_block(b, m, x, **kw):
    """MixConvBlock.forward (mixnet.py:151-157) with MixConv.forward (mixnet.py:76-80): split the input channels, one conv per
    part (1x1 parts, or depthwise parts with kernels 3, 5, 7, ...), concatenate, normalise, activate.  Every part reads and writes
    a channel slice of one buffer (parts padded to 8 channels, TRef.cmap); the shared BatchNorm is folded part by part."""
    if kw.get("residual") is not None or kw.get("post_act") is not None or kw.get("out") is not None:
        raise NotImplementedError("MixConvBlock takes no fused residual / concat slice")
    mc = m.conv
    if mc.axis != 1:
        raise NotImplementedError("MixConv splits along the channel axis")
    convs = list(mc.children())
    sizes_in = list(mc.splitted_in_channels)
    sizes_out = [c.out_channels for c in convs]
    in_map, in_storage, in_offs = _seg_layout(sizes_in)
    x = b.relayout(x, in_map, in_storage)
    out_map, out_storage, out_offs = _seg_layout(sizes_out)
    c0 = convs[0]
    Ho = (x.H + 2 * _one(c0.padding) - _one(c0.dilation) * (c0.kernel_size[0] - 1) - 1) // _one(c0.stride) + 1
    Wo = (x.W + 2 * _one(c0.padding) - _one(c0.dilation) * (c0.kernel_size[1] - 1) - 1) // _one(c0.stride) + 1
    cat = b.new(x.N, Ho, Wo, out_storage)
    act = act_code(m.activ) if m.activate else ACT_NONE
    if m.activate and _standalone_act(m.activ):
        raise NotImplementedError("MixConvBlock with a stand-alone activation")
    lo = 0
    for conv, n_in, n_out, oi, oo in zip(convs, sizes_in, sizes_out, in_offs, out_offs):
        xi = Builder.view(x, oi, _rup(n_in, 8))
        xi.cmap = None if n_in % 8 == 0 else list(range(n_in))
        bn = _bn_slice(m.bn, lo, lo + n_out) if m.normalize else None
        y = b.conv(xi, conv, bn, act, out=Builder.view(cat, oo, _rup(n_out, 8)),
                   act_a=act_param(m.activ) if m.activate else 0.0)
        if y.buf is not cat.buf:
            raise NotImplementedError("MixConv parts must write into the concat buffer")
        lo += n_out
    cat.cmap = out_map
    return cat


@lowers("MixUnit")
def _lower_mix_unit(b, m, x, **kw):
    """MixUnit.forward (mixnet.py:282-293): [expansion] -> (mixed) depthwise -> [SE] -> (mixed) 1x1 linear (+x)."""
    y = lower(b, m.exp_conv, x) if m.use_exp_conv else x
    y = lower(b, m.conv1, y)
    if m.use_se:
        y = lower(b, m.se, y)
    if not m.residual:
        return lower(b, m.conv2, y)
    if type(m.conv2).__name__ == "ConvBlock" and x.cmap is None:
        return lower(b, m.conv2, y, residual=x, post_act=None)       # the add rides on the 1x1 conv's epilogue
    y = lower(b, m.conv2, y)
    return b.add_act(b.relayout(y, x.cmap, x.C), x, ACT_NONE)


@lowers("MixInitBlock")
def _lower_mix_init(b, m, x, **kw):
    """MixInitBlock.forward (mixnet.py:326-329)."""
    return lower(b, m.conv2, lower(b, m.conv1, x))


@lowers("DarkUnit")
def _lower_dark_unit(b, m, x, **kw):
    """DarkUnit.forward (darknet53.py:45-49): conv1x1 -> conv3x3 (each with its LeakyReLU) + x, nothing after the add."""
    return lower(b, m.conv2, lower(b, m.conv1, x), residual=x, post_act=ACT_NONE)


def _tf_pad(m, x: TRef, kernel_size: int, stride: int = 1, dilation: int = 1):
    """calc_tf_padding (efficientnet.py:27-55) for a tf_mode unit, None otherwise.  The reference hands the tuple
    (pad_h//2, pad_h - pad_h//2, pad_w//2, pad_w - pad_w//2) to F.pad, whose order is (left, right, top, bottom): the
    HEIGHT amounts land on the width axis and vice versa - reproduced as is (identical for square maps)."""
    if not getattr(m, "tf_mode", False):
        return None
    oh, ow = -(-x.H // stride), -(-x.W // stride)
    ph = max((oh - 1) * stride + (kernel_size - 1) * dilation + 1 - x.H, 0)
    pw = max((ow - 1) * stride + (kernel_size - 1) * dilation + 1 - x.W, 0)
    return (ph // 2, ph - ph // 2, pw // 2, pw - pw // 2)


@lowers("EffiInitBlock")
def _lower_effi_init(b, m, x, **kw):
    """EffiInitBlock.forward (efficientnet.py:235-239)."""
    return lower(b, m.conv, x, pad_lrtb=_tf_pad(m, x, 3, 2))


@lowers("EffiDwsConvUnit")
def _lower_effi_dws(b, m, x, **kw):
    """EffiDwsConvUnit.forward (efficientnet.py:105-115): dw3x3 -> SE -> 1x1 linear (+x); the add rides on the 1x1."""
    y = lower(b, m.se, lower(b, m.dw_conv, x, pad_lrtb=_tf_pad(m, x, 3)))
    return lower(b, m.pw_conv, y, residual=x if m.residual else None, post_act=None)


@lowers("EffiInvResUnit")
def _lower_effi_invres(b, m, x, **kw):
    """EffiInvResUnit.forward (efficientnet.py:185-197): 1x1 expand -> dw kxk -> [SE] -> 1x1 linear (+x)."""
    y1 = lower(b, m.conv1, x)
    y = lower(b, m.conv2, y1, pad_lrtb=_tf_pad(m, y1, m.kernel_size, m.stride))
    if m.use_se:
        y = lower(b, m.se, y)
    return lower(b, m.conv3, y, residual=x if m.residual else None, post_act=None)


@lowers("EffiEdgeResUnit")
def _lower_effi_edge_unit(b, m, x, **kw):
    """EffiEdgeResUnit.forward (efficientnetedge.py:77-86): conv3x3 expansion -> [SE] -> 1x1 linear (+x); the add rides on the 1x1."""
    y = lower(b, m.conv1, x)
    if m.use_se:
        y = lower(b, m.se, y)
    return lower(b, m.conv2, y, residual=x if m.residual else None, post_act=None)


@lowers("MobileNetV3Unit")
def _lower_mnv3_unit(b, m, x, **kw):
    """MobileNetV3Unit.forward (mobilenetv3.py:82-93): [1x1 expand] -> dw -> [SE] -> 1x1 linear (+x)."""
    y = lower(b, m.exp_conv, x) if m.use_exp_conv else x
    y = lower(b, m.conv1, y)
    if m.use_se:
        y = lower(b, m.se, y)
    return lower(b, m.conv2, y, residual=x if m.residual else None, post_act=None)


@lowers("DwsExpSEResUnit")
def _lower_dws_exp_se_res(b, m, x, **kw):
    """DwsExpSEResUnit.forward (mnasnet.py:77-88): [1x1 expand] -> dw -> [SE] -> 1x1 linear (+x)."""
    y = lower(b, m.exp_conv, x) if m.use_exp_conv else x
    y = lower(b, m.dw_conv, y)
    if m.use_se:
        y = lower(b, m.se, y)
    return lower(b, m.pw_conv, y, residual=x if m.residual else None, post_act=None)


@lowers("ProxylessUnit")
def _lower_proxyless_unit(b, m, x, **kw):
    """ProxylessUnit.forward (proxylessnas.py:114-123) with ProxylessBlock.forward (:64-70): identity | body | x + body(x)."""
    if not m.residual:
        return x
    blk = m.body
    if blk.use_bc:
        return _exp_dw_pw(b, blk.bc_conv, blk.dw_conv, blk.pw_conv, x, residual=x if m.shortcut else None, post_act=None)
    return _dw_then_pw(b, blk.dw_conv, blk.pw_conv, x, residual=x if m.shortcut else None, post_act=None)


@lowers("FBNetUnit", "SPNASUnit")
def _lower_fbnet_unit(b, m, x, **kw):
    """FBNetUnit.forward (fbnet.py:77-87) == SPNASUnit.forward (spnasnet.py:72-82): [1x1 expand] -> dw -> 1x1 (+x)."""
    if m.use_exp_conv:
        return _exp_dw_pw(b, m.exp_conv, m.conv1, m.conv2, x, residual=x if m.residual else None, post_act=None)
    return _dw_then_pw(b, m.conv1, m.conv2, x, residual=x if m.residual else None, post_act=None)


@lowers("MnasInitBlock", "MnasFinalBlock", "FBNetInitBlock", "SPNASInitBlock", "SPNASFinalBlock")
def _lower_mnas_edge(b, m, x, **kw):
    """MnasInitBlock.forward / MnasFinalBlock.forward (mnasnet.py:121-124, 157-160): conv1 then conv2."""
    return lower(b, m.conv2, lower(b, m.conv1, x))


@lowers("MobileNetV3FinalBlock")
def _lower_mnv3_final(b, m, x, **kw):
    """MobileNetV3FinalBlock.forward (mobilenetv3.py:127-131)."""
    y = lower(b, m.conv, x)
    return lower(b, m.se, y) if m.use_se else y


@lowers("MobileNetV3Classifier")
def _lower_mnv3_classifier(b, m, x, **kw):
    """MobileNetV3Classifier.forward (mobilenetv3.py:168-174): the h-swish rides on conv1's epilogue, fp32 logits."""
    y = b.conv(x, m.conv1, None, act_code(m.activ))
    return b.conv(y, m.conv2, None, ACT_NONE, out_f32=True)


@lowers("MobileNetV3")
def _lower_mobilenetv3(b, m, x, **kw):
    """MobileNetV3.forward (mobilenetv3.py:277-281): features -> classifier convs on the 1x1 map -> view."""
    return _flat(lower(b, m.output, lower(b, m.features, x)))


@lowers("MultiOutputSequential")
def _lower_multi_output(b, m, x, **kw):
    """MultiOutputSequential.forward (arch.py:332-347)."""
    outs = []
    for child in m.children():
        x = lower(b, child, x)
        if getattr(child, "do_output", False):
            outs.append(x)
        elif getattr(child, "do_output2", False):
            raise NotImplementedError("do_output2 children are outside the B200 eval path")
    if m.multi_output:
        return [x] + outs if m.return_last else outs
    if m.dual_output:
        return x, outs
    return x


@lowers("Concurrent")
def _lower_concurrent(b, m, x, **kw):
    """Concurrent.forward (arch.py:84-95), merge_type "cat": branches write straight into channel slices."""
    if m.merge_type != "cat" or m.axis != 1:
        raise NotImplementedError("only Concurrent(cat, axis=1) is supported")
    branches = list(m.children())
    widths = [_branch_width(br, x.C) for br in branches]
    Ho, Wo = x.H, x.W
    cat = b.new(x.N, Ho, Wo, sum(widths))
    off = 0
    for br, wd in zip(branches, widths):
        if type(br).__name__ == "Identity":
            # PyramidPooling's first branch (pspnet.py:109): the input itself becomes a channel slice of the concat
            # buffer - a same-size align_corners resample is an exact copy (weights 1 / 0) into the pitched slice
            y = b.bilinear(x, x.H, x.W, out=Builder.view(cat, off, wd))
            off += wd
            continue
        y = lower(b, br, x, out=Builder.view(cat, off, wd))
        if y.buf is not cat.buf:
            raise NotImplementedError(f"branch {type(br).__name__} cannot write into a concat slice")
        off += wd
    return cat


def _branch_width(br: nn.Module, in_channels: int | None = None) -> int:
    convs = [mod for mod in br.modules() if isinstance(mod, nn.Conv2d)]
    if not convs:
        if type(br).__name__ == "Identity" and in_channels is not None:
            return in_channels
        raise NotImplementedError(f"cannot infer the width of branch {type(br).__name__}")
    return convs[-1].out_channels


@lowers("ASPPAvgBranch")
def _lower_aspp_avg(b, m, x, out=None, **kw):
    """ASPPAvgBranch.forward (deeplabv3.py:80-87): global mean -> 1x1 ConvBlock -> bilinear broadcast."""
    size = m.upscale_out_size if m.upscale_out_size is not None else (x.H, x.W)
    y = lower(b, m.conv, b.gap(x))
    return b.bilinear(y, size[0], size[1], out=out)


@lowers("PyramidPoolingBranch")
def _lower_pyramid_branch(b, m, x, out=None, **kw):
    """PyramidPoolingBranch.forward (pspnet.py:71-75): adaptive pool -> 1x1 ConvBlock -> bilinear back to the map size."""
    size = m.upscale_out_size if m.upscale_out_size is not None else (x.H, x.W)
    y = lower(b, m.conv, lower(b, m.pool, x))
    return b.bilinear(y, size[0], size[1], out=out)


@lowers("PyramidPooling")
def _lower_pyramid_pooling(b, m, x, **kw):
    return lower(b, m.branches, x)


@lowers("AtrousSpatialPyramidPooling")
def _lower_aspp(b, m, x, **kw):
    return lower(b, m.dropout, lower(b, m.conv, lower(b, m.branches, x)))


@lowers("DeepLabv3FinalBlock", "FCNFinalBlock", "PSPFinalBlock")
def _lower_deeplab_final(b, m, x, out_size=None, **kw):
    """DeepLabv3FinalBlock.forward (deeplabv3.py:49-54) == FCNFinalBlock.forward (fcn8sd.py:47-52); the result is the
    fp32 NCHW tensor the reference returns."""
    y = lower(b, m.conv2, lower(b, m.dropout, lower(b, m.conv1, x)))
    return b.bilinear(y, out_size[0], out_size[1], nchw_f32=True)


@lowers("DeepLabv3")
def _lower_deeplab(b, m, x, **kw):
    """DeepLabv3.forward (deeplabv3.py:199-208)."""
    in_size = m.in_size if m.fixed_size else (x.H, x.W)
    feats = lower(b, m.backbone, x)
    x4, x3 = feats[0], feats[1]
    y = lower(b, m.final_block, lower(b, m.pool, x4), out_size=in_size)
    if m.aux:
        return y, lower(b, m.aux_block, x3, out_size=in_size)
    return y


@lowers("FCN8sd", "PSPNet")
def _lower_fcn8sd(b, m, x, **kw):
    """FCN8sd.forward (fcn8sd.py:112-120); PSPNet.forward (pspnet.py:196-205) adds the pyramid pooling module."""
    in_size = m.in_size if m.fixed_size else (x.H, x.W)
    feats = lower(b, m.backbone, x)
    x4, x3 = feats[0], feats[1]
    if hasattr(m, "pool"):
        x4 = lower(b, m.pool, x4)
    y = lower(b, m.final_block, x4, out_size=in_size)
    if m.aux:
        return y, lower(b, m.aux_block, x3, out_size=in_size)
    return y


def _flat(t: TRef) -> TRef:
    t.flat = True
    return t


@lowers("ResNet", "SEResNeXt", "SEResNet", "ResNeXt", "MobileNet", "EfficientNet", "MnasNet", "FBNet", "SPNASNet", "SENet", "ProxylessNAS",
        "PreResNet", "DarkNet53", "MixNet", "EfficientNetEdge")
def _lower_classifier(b, m, x, **kw):
    """features -> view(N,-1) -> [Dropout ->] Linear (resnet.py:333-337, seresnext.py:136-140, efficientnet.py:354-358)."""
    return _flat(lower(b, m.output, lower(b, m.features, x)))


@lowers("MobileNetV2")
def _lower_mobilenetv2(b, m, x, **kw):
    """features -> bare conv1x1 classifier on the 1x1 map -> view (mobilenetv2.py:152-156)."""
    return _flat(lower(b, m.output, lower(b, m.features, x), out_f32=True))


@lowers("ResNetD")
def _lower_resnetd(b, m, x, **kw):
    """ResNetD.forward (resnetd.py:98-106)."""
    outs = lower(b, m.features, x)
    if not isinstance(outs, list):
        outs = [outs]
    logits = _flat(lower(b, m.output, outs[0]))
    return [logits] + outs[1:] if m.multi_output else logits


# ---------------------------------------------------------------------------------------------------------------
# compiled module
# ---------------------------------------------------------------------------------------------------------------
def _assign_offsets(bufs: list[Buf]) -> int:
    """Greedy lowest-offset placement with lifetime-overlap checks; returns the arena size."""
    placed: list[Buf] = []
    top = 0
    for buf in sorted(bufs, key=lambda z: (z.first, -z.nbytes)):
        size = _rup(buf.nbytes, _ALIGN)
        lo, hi = buf.first, (1 << 60) if buf.pinned else max(buf.last, buf.first)
        conflicts = sorted(
            ((p.offset, p.offset + _rup(p.nbytes, _ALIGN)) for p in placed
             if not (((1 << 60) if p.pinned else max(p.last, p.first)) < lo or p.first > hi)),
            key=lambda iv: iv[0])
        off = 0
        for s, e in conflicts:
            if off + size <= s:
                break
            off = max(off, e)
        buf.offset = off
        placed.append(buf)
        top = max(top, off + size)
    return top


def weights_signature(module: nn.Module) -> tuple:
    return tuple((id(t), t._version) for t in list(module.parameters()) + list(module.buffers()))


class CompiledModule:
    """A module tree compiled for one input shape / tier / device."""

    def __init__(self, module: nn.Module, in_shape: tuple[int, int, int, int], dtype="bf16",
                 device: torch.device | str | None = None, graph: bool = False, lower_kwargs: dict | None = None,
                 alias_outputs: bool = False, input_affine: tuple | None = None):
        """`input_affine=(scale, bias)` (per-channel sequences) applies x * scale[c] + bias[c] inside the ingest kernel: a
        uint8 pipeline passes (1 / (255 * std), -mean / std) and hands the raw uint8 NCHW batch to forward().

        `alias_outputs=True` returns the SAME tensors on every call (views of plan-owned memory, overwritten by the
        next forward): a zero-copy opt-in for pipelines that consume a result before the next call.  The default follows
        the reference's contract (resnet.py:333-337): every call returns freshly allocated, caller-owned
        tensors."""
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("pytorchcv_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.dtype = dtype_code(dtype)
        self.in_shape = tuple(int(v) for v in in_shape)
        self.use_graph = graph
        self.alias_outputs = bool(alias_outputs)
        self._affine = None
        if input_affine is not None:
            sc, bi = ([float(v) for v in t] for t in input_affine)
            if len(sc) != in_shape[1] or len(bi) != in_shape[1] or in_shape[1] > 4:
                raise ValueError("input_affine needs one scale and one bias per image channel (<= 4 channels)")
            self._affine = ((C.c_float * len(sc))(*sc), (C.c_float * len(bi))(*bi))
        self.signature = weights_signature(module)
        N, Cin, H, W = self.in_shape

        for allow_s2d, allow_pool in ((True, True), (True, False), (False, False)):
            b = Builder(self.dtype, self.device, allow_s2d_stem=allow_s2d, allow_stem_pool=allow_pool)
            x = b.new(N, H, W, _rup(Cin, 8))
            x.buf.first = -1
            x.buf.pinned = True
            b.input_tref, b.image_channels = x, Cin
            try:
                result = lower(b, module, x, **(lower_kwargs or {}))
                break
            except (_NoStemPool, _NoS2dStem):
                continue   # retry order: drop the pool fusion first, then the s2d stem
        self._stem_k = b.stem["k"] if b.stem else 0
        self._flops_override = dict(b.flops_override)
        self._in = b.stem["tref"] if b.stem else x
        self._in_channels = Cin
        self._structure, trefs = _flatten(result)
        outs = []
        for t in trefs:
            if t.layout == "nhwc" and not (t.flat and t.dtype == F32):
                flat = t.flat
                t = b.egress(t)
                t.flat = flat
            if t.tail is None:
                t.buf.pinned = True
                t.buf.first = -1   # stable from one forward to the next (never part of another tensor's slot)
            outs.append(t)
        self._outs = outs

        arena_bytes = _assign_offsets(b.bufs)
        with torch.cuda.device(self.device):
            self.arena = torch.zeros(arena_bytes + _ALIGN, dtype=torch.uint8, device=self.device)
            self.weights = torch.zeros(b.weight_bytes + _ALIGN, dtype=torch.uint8, device=self.device)
            abase = _rup(self.arena.data_ptr(), _ALIGN)
            wbase = _rup(self.weights.data_ptr(), _ALIGN)
            self._abase = abase

            def ptr(t: TRef) -> int:
                return abase + t.buf.offset + t.byte_off

            def wptr(off: int) -> int:
                return wbase + off

            self._pack_weights(b, wptr)
            handle = C.c_void_p()
            _lib.call("pcv_plan_create", C.byref(handle))
            self._plan = handle
            for emit in b.ops:
                emit(self._plan, ptr, wptr)
            torch.cuda.synchronize(self.device)
            self._ptr = ptr
        self._in_ptr = ptr(self._in)
        self.arena_bytes = arena_bytes
        self.weight_bytes = b.weight_bytes
        self.num_ops = lib.pcv_plan_num_ops(self._plan)
        self.num_launches = lib.pcv_plan_num_launches(self._plan) + 1 + sum(t.tail is not None for t in outs)  # + ingest + edge ops
        self._arena_views = [self._tensor_of(t) if t.tail is None else None for t in outs]
        self._static_outs = None   # alias_outputs: edge tensors allocated once
        self._graph_stream = torch.cuda.Stream(device=self.device) if graph else None

    # -- weights ------------------------------------------------------------------------------------------------
    def _dev_f32(self, t: torch.Tensor | None):
        if t is None:
            return None
        return t.detach().to(device=self.device, dtype=torch.float32).contiguous()

    def _pack_weights(self, b: Builder, wptr) -> None:
        keep = []
        stream = torch.cuda.current_stream(self.device).cuda_stream
        for kind, payload in b.weight_jobs:
            if kind == "raw":
                t, off = payload
                src = self._dev_f32(t)
                keep.append(src)
                rel = wptr(off) - self.weights.data_ptr()
                self.weights[rel:rel + src.numel() * 4].copy_(src.view(-1).view(torch.uint8))
                continue
            if kind == "conv_mapped":
                self._pack_mapped(payload, keep, stream, wptr)
                continue
            if kind == "conv_dual":
                self._pack_dual(payload, keep, stream, wptr)
                continue
            if kind == "conv_s2d":
                d, conv, bn, k, w_off, b_off = payload
                w0 = self._dev_f32(conv.weight)
                w = torch.empty((conv.out_channels, d.Cin, d.kh, 1), dtype=torch.float32, device=self.device)
                _lib.call("pcv_stem_s2d_weights", conv.out_channels, conv.in_channels, k, w0.data_ptr(), w.data_ptr(),
                          stream)
                keep.append(w0)
                pad_cin = 0
            else:
                d, conv, bn, pad_cin, w_off, b_off = payload
                w = self._dev_f32(conv.weight)
            if pad_cin:
                w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, pad_cin)).contiguous()
            cb = self._dev_f32(conv.bias)
            if bn is not None:
                g = self._dev_f32(bn.weight) if bn.weight is not None else torch.ones_like(self._dev_f32(bn.running_var))
                be = self._dev_f32(bn.bias) if bn.bias is not None else torch.zeros_like(g)
                mu, var, eps = self._dev_f32(bn.running_mean), self._dev_f32(bn.running_var), float(bn.eps)
            else:
                g = be = mu = var = None
                eps = 0.0
            keep += [w, cb, g, be, mu, var]
            _lib.call("pcv_pack_conv_weights", C.byref(d), self.dtype, w.data_ptr(),
                      cb.data_ptr() if cb is not None else None,
                      g.data_ptr() if g is not None else None, be.data_ptr() if be is not None else None,
                      mu.data_ptr() if mu is not None else None, var.data_ptr() if var is not None else None,
                      eps, wptr(w_off), wptr(b_off), stream)
        torch.cuda.synchronize(self.device)
        del keep

    def _pack_one(self, d, conv, bn, w_ptr: int, b_ptr: int, keep, stream) -> None:
        """pcv_pack_conv_weights of one ConvBlock's conv (+ BatchNorm, folded by the library) at explicit device addresses."""
        w, cb = self._dev_f32(conv.weight), self._dev_f32(conv.bias)
        g = be = mu = var = None
        eps = 0.0
        if bn is not None:
            g = self._dev_f32(bn.weight) if bn.weight is not None else torch.ones_like(self._dev_f32(bn.running_var))
            be = self._dev_f32(bn.bias) if bn.bias is not None else torch.zeros_like(g)
            mu, var, eps = self._dev_f32(bn.running_mean), self._dev_f32(bn.running_var), float(bn.eps)
        keep += [w, cb, g, be, mu, var]
        _lib.call("pcv_pack_conv_weights", C.byref(d), self.dtype, w.data_ptr(), cb.data_ptr() if cb is not None else None,
                  g.data_ptr() if g is not None else None, be.data_ptr() if be is not None else None,
                  mu.data_ptr() if mu is not None else None, var.data_ptr() if var is not None else None, eps, w_ptr, b_ptr,
                  stream)

    def _pack_dual(self, payload, keep, stream, wptr) -> None:
        """Weights of Builder.conv_dual: both 1x1 convs packed as usual, then joined row by row along K; biases summed."""
        d, cb, d2, cb2, sizes, w_off, b_off, b2_off = payload
        parts = []
        for dd, m, (wn, bn_) in ((d, cb, sizes[0]), (d2, cb2, sizes[1])):
            wt = torch.empty(wn, dtype=torch.uint8, device=self.device)
            bt = torch.empty(bn_ // 4, dtype=torch.float32, device=self.device)
            self._pack_one(dd, m.conv, m.bn if m.normalize else None, wt.data_ptr(), bt.data_ptr(), keep, stream)
            parts.append((wt.view(d.Cout, -1), bt))
        joined = torch.cat([parts[0][0], parts[1][0]], dim=1).contiguous().view(-1)
        blobs = [(w_off, joined)]
        if b2_off is None:
            blobs.append((b_off, (parts[0][1] + parts[1][1]).view(torch.uint8).view(-1)))
        else:
            blobs += [(b_off, parts[0][1].view(torch.uint8).view(-1)), (b2_off, parts[1][1].view(torch.uint8).view(-1))]
        for off, t in blobs:
            keep.append(t)
            rel = wptr(off) - self.weights.data_ptr()
            self.weights[rel:rel + t.numel()].copy_(t)

    def _pack_mapped(self, payload, keep, stream, wptr) -> None:
        """Weights of a conv with virtual channel padding (Builder._conv_mapped): scatter into the storage channel space."""
        d, conv, bn, in_idx, cs_in, out_idx, cs_out, depthwise, w_off, b_off = payload
        dev = self.device
        w0 = self._dev_f32(conv.weight)
        oi = torch.tensor(out_idx, dtype=torch.long, device=dev)
        ii = torch.tensor(in_idx, dtype=torch.long, device=dev)
        kh, kw = w0.shape[2], w0.shape[3]
        if depthwise:
            w = torch.zeros((cs_out, 1, kh, kw), dtype=torch.float32, device=dev)
            w[oi] = w0
        else:
            w = torch.zeros((cs_out, cs_in, kh, kw), dtype=torch.float32, device=dev)
            w[oi[:, None], ii[None, :]] = w0
        cb = None
        if conv.bias is not None:
            cb = torch.zeros(cs_out, dtype=torch.float32, device=dev)
            cb[oi] = self._dev_f32(conv.bias)
        g = be = mu = var = None
        eps = 0.0
        if bn is not None:
            # padding channels: identity BatchNorm on a zero weight row and a zero bias -> exact zeros
            g = torch.ones(cs_out, dtype=torch.float32, device=dev)
            be = torch.zeros(cs_out, dtype=torch.float32, device=dev)
            mu = torch.zeros(cs_out, dtype=torch.float32, device=dev)
            var = torch.ones(cs_out, dtype=torch.float32, device=dev)
            if bn.weight is not None:
                g[oi] = self._dev_f32(bn.weight)
            if bn.bias is not None:
                be[oi] = self._dev_f32(bn.bias)
            mu[oi] = self._dev_f32(bn.running_mean)
            var[oi] = self._dev_f32(bn.running_var)
            eps = float(bn.eps)
        keep += [w0, w, cb, g, be, mu, var]
        _lib.call("pcv_pack_conv_weights", C.byref(d), self.dtype, w.data_ptr(),
                  cb.data_ptr() if cb is not None else None,
                  g.data_ptr() if g is not None else None, be.data_ptr() if be is not None else None,
                  mu.data_ptr() if mu is not None else None, var.data_ptr() if var is not None else None,
                  eps, wptr(w_off), wptr(b_off), stream)

    # -- outputs ------------------------------------------------------------------------------------------------
    def _tensor_of(self, t: TRef) -> torch.Tensor:
        if t.dtype != F32:
            raise AssertionError("network outputs are fp32")
        start = (self._abase - self.arena.data_ptr()) + t.buf.offset
        if t.layout == "nchw":
            n = t.N * t.C * t.H * t.W
            out = self.arena[start:start + 4 * n].view(torch.float32).view(t.N, t.C, t.H, t.W)
        else:  # flat fp32 logits written NHWC with H = W = 1
            n = t.N * t.pitch
            out = self.arena[start:start + 4 * n].view(torch.float32).view(t.N, t.pitch)[:, :t.C]
        if t.flat:
            out = out.reshape(t.N, -1)
        return out

    # -- run ----------------------------------------------------------------------------------------------------
    def __call__(self, x: torch.Tensor):
        if tuple(x.shape) != self.in_shape:
            raise ValueError(f"compiled for input {self.in_shape}, got {tuple(x.shape)}")
        if x.device != self.device:
            raise ValueError(f"compiled for {self.device}, input is on {x.device}: there is no CPU fallback")
        if x.dtype not in _IMG_TYPES:
            x = x.float()
        if not x.is_contiguous():
            x = x.contiguous()
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream(self.device)
            if self.use_graph:
                gs = self._graph_stream
                gs.wait_stream(cur)
                self._ingest(x, gs.cuda_stream)
                _lib.call("pcv_plan_graph_launch", self._plan, gs.cuda_stream)
                cur.wait_stream(gs)
                x.record_stream(gs)
            else:
                self._ingest(x, cur.cuda_stream)
                _lib.call("pcv_plan_run", self._plan, cur.cuda_stream)
            return _unflatten(self._structure, self._outputs(cur))

    def _outputs(self, cur) -> list:
        """The tensors of this call (on stream `cur`, after the plan): edge ops write straight into fresh tensors, logits
        are copied out of the arena (<= 1 MB).  With alias_outputs the same plan-owned tensors are returned every time."""
        if self.alias_outputs and self._static_outs is None:
            self._static_outs = [v if t.tail is None else self._shape(torch.empty(
                (t.N, t.C, t.H, t.W), dtype=torch.float32, device=self.device), t)
                for t, v in zip(self._outs, self._arena_views)]
        res = []
        for i, (t, view) in enumerate(zip(self._outs, self._arena_views)):
            if t.tail is None:
                res.append(view if self.alias_outputs else view.clone())
                continue
            o = (self._static_outs[i] if self.alias_outputs
                 else self._shape(torch.empty((t.N, t.C, t.H, t.W), dtype=torch.float32, device=self.device), t))
            t.tail["launch"](self._ptr, o.data_ptr(), cur.cuda_stream)
            res.append(o)
        return res

    @staticmethod
    def _shape(o: torch.Tensor, t: TRef) -> torch.Tensor:
        return o.view(t.N, -1) if t.flat else o

    def _ingest(self, x: torch.Tensor, stream: int) -> None:
        """The network edge: NCHW image (fp32 as in the reference, or a bf16 / fp16 / uint8 copy of it) -> NHWC
        (channel-padded) or, for a strided stem, space-to-depth."""
        N, Cin, H, W = self.in_shape
        sc, bi = self._affine if self._affine is not None else (None, None)
        if self._stem_k:
            _lib.call("pcv_stem_s2d_ingest_ex", None, self.dtype, _IMG_TYPES[x.dtype], N, Cin, H, W, self._stem_k,
                      x.data_ptr(), sc, bi, self._in_ptr, stream)
        else:
            _lib.call("pcv_nchw_to_nhwc_ex", None, self.dtype, _IMG_TYPES[x.dtype], N, Cin, H, W, x.data_ptr(), sc, bi,
                      self._in_ptr, self._in.pitch, stream)

    def profile(self) -> list[tuple[str, float, float, float]]:
        """[(op name, ms, algorithmic FLOPs, algorithmic bytes)] for one eager pass (synchronises); the per-call edge ops
        (fp32 NCHW outputs written after the plan) are timed the same way and appended."""
        lib = _lib.load()
        n = self.num_ops
        ms = (C.c_float * n)()
        rows = []
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream(self.device)
            _lib.call("pcv_plan_profile", self._plan, cur.cuda_stream, ms, n)
            for i in range(n):
                fl, by = C.c_double(), C.c_double()
                _lib.call("pcv_plan_op_cost", self._plan, i, C.byref(fl), C.byref(by))
                rows.append((lib.pcv_plan_op_name(self._plan, i).decode(), float(ms[i]),
                             self._flops_override.get(i, fl.value), by.value))
            for t in self._outs:
                if t.tail is None:
                    continue
                o = torch.empty((t.N, t.C, t.H, t.W), dtype=torch.float32, device=self.device)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(cur)
                t.tail["launch"](self._ptr, o.data_ptr(), cur.cuda_stream)
                e1.record(cur)
                e1.synchronize()
                rows.append((t.tail["name"], e0.elapsed_time(e1), 0.0, t.tail["bytes"]))
        return rows

    def __del__(self):
        plan = getattr(self, "_plan", None)
        if plan is not None and _lib._lib is not None:
            try:
                _lib._lib.pcv_plan_destroy(plan)
            except Exception:
                pass
            self._plan = None


_IMG_TYPES = {torch.float32: _lib.IMG_F32, torch.bfloat16: _lib.IMG_BF16, torch.float16: _lib.IMG_F16,
              torch.uint8: _lib.IMG_U8}


def _flatten(obj):
    if isinstance(obj, TRef):
        return None, [obj]
    if isinstance(obj, (list, tuple)):
        specs, flat = [], []
        for o in obj:
            s, f = _flatten(o)
            specs.append((s, len(f)))
            flat += f
        return (type(obj), specs), flat
    raise TypeError(f"unexpected lowering result {type(obj)}")


def _unflatten(spec, flat):
    if spec is None:
        return flat.pop(0)
    typ, specs = spec
    return typ(_unflatten(s, flat) for s, _ in specs)


# ---------------------------------------------------------------------------------------------------------------
# per-module cache used by the mirror modules' forward() and by accelerate()
# ---------------------------------------------------------------------------------------------------------------
_DEFAULT = {"dtype": "bf16", "graph": True}   # mirror modules: compile on first call, CUDA-graph replay after
# compiled plans per module instance; kept OUT of module.__dict__ so deepcopy / pickle / state_dict never see them
_CACHES: "weakref.WeakKeyDictionary[nn.Module, dict]" = weakref.WeakKeyDictionary()


def plan_cache(module: nn.Module) -> dict:
    return _CACHES.setdefault(module, {})


def set_default_precision(dtype: str) -> None:
    """Tier used by modules that were not explicitly accelerate()d: "bf16" (default) or "fp32"."""
    dtype_code(dtype)
    _DEFAULT["dtype"] = dtype


def _affine_key(a):
    return None if a is None else (tuple(float(v) for v in a[0]), tuple(float(v) for v in a[1]))


def set_default_graph(enabled: bool) -> None:
    """Whether modules that were not explicitly accelerate()d replay their plan from a CUDA graph (default) or launch it
    kernel by kernel."""
    _DEFAULT["graph"] = bool(enabled)


def run_module(module: nn.Module, x: torch.Tensor, dtype=None, graph=None, check_weights: bool = True,
               alias_outputs: bool = False, input_affine: tuple | None = None, **lower_kwargs):
    """forward() of every mirror block/net: compile on first use (per input shape & tier), then run the plan."""
    if not isinstance(x, torch.Tensor) or x.dim() != 4:
        raise ValueError("expected an NCHW tensor")
    if not x.is_cuda:
        raise RuntimeError("pytorchcv_b200 runs on CUDA (sm_100a) only; there is no CPU fallback — move the input "
                           "to the GPU or use the reference package on CPU")
    dtype = _DEFAULT["dtype"] if dtype is None else dtype
    graph = _DEFAULT["graph"] if graph is None else graph
    cache = plan_cache(module)
    key = (tuple(x.shape), dtype_code(dtype), x.device.index, bool(graph), bool(alias_outputs),
           _affine_key(input_affine), tuple(sorted(lower_kwargs.items())))
    cm = cache.get(key)
    if cm is not None and check_weights and cm.signature != weights_signature(module):
        cm = None  # parameters were replaced or modified in place (load_state_dict, .to(), optimizer step)
    if cm is None:
        cm = CompiledModule(module, tuple(x.shape), dtype=dtype, device=x.device, graph=graph,
                            lower_kwargs=lower_kwargs, alias_outputs=alias_outputs, input_affine=input_affine)
        cache[key] = cm
    return cm(x)


def invalidate(module: nn.Module) -> None:
    """Drop every compiled plan cached on `module` (and its sub-modules)."""
    for mod in module.modules():
        _CACHES.pop(mod, None)


class Accelerated(nn.Module):
    """Drop-in wrapper returned by accelerate(): same call signature as the wrapped reference module."""

    def __init__(self, net: nn.Module, dtype="bf16", graph: bool = False, check_weights: bool = True,
                 alias_outputs: bool = False, input_affine: tuple | None = None):
        super().__init__()
        self.net = net
        self._dtype, self._graph, self._check, self._alias = dtype, graph, check_weights, alias_outputs
        self._affine = input_affine

    def forward(self, x):
        return run_module(self.net, x, dtype=self._dtype, graph=self._graph, check_weights=self._check,
                          alias_outputs=self._alias, input_affine=self._affine)

    def compiled(self, x: torch.Tensor) -> CompiledModule:
        self.forward(x)
        key = (tuple(x.shape), dtype_code(self._dtype), x.device.index, bool(self._graph), bool(self._alias),
               _affine_key(self._affine), ())
        return plan_cache(self.net)[key]


def accelerate(net: nn.Module, dtype="bf16", graph: bool = False, check_weights: bool = True,
               alias_outputs: bool = False, input_affine: tuple | None = None) -> Accelerated:
    """Compile an eval-mode pytorchcv network (reference or mirror modules) for the B200 path.

    Opt-in per model instance (the reference's own _test()s call .backward() on eval nets, so a
    global monkey-patch of ConvBlock.forward would break them)."""
    if net.training:
        raise RuntimeError("pytorchcv_b200 is an eval-mode path: call net.eval() first")
    dtype_code(dtype)
    return Accelerated(net, dtype=dtype, graph=graph, check_weights=check_weights, alias_outputs=alias_outputs,
                       input_affine=input_affine)
