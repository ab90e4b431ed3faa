"""ctypes binding of libpcv_b200.so (declared in include/pcv_b200.h).

The library is the product: if it is missing or cannot be loaded this module raises — there is no Python / CPU
fallback for any op ("Unsupported module patterns are rejected ... no CPU fallback").
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PCV_B200_LIB") or os.path.join(_HERE, "libpcv_b200.so")   # env: A/B a second build

BF16, F32, F16 = 0, 1, 2                      # pcv_dtype
IMG_F32, IMG_BF16, IMG_F16, IMG_U8 = 0, 1, 2, 3  # pcv_image_type
ACT_NONE, ACT_RELU, ACT_RELU6, ACT_SIGMOID, ACT_SWISH, ACT_HSWISH, ACT_HSIGMOID, ACT_LEAKY_RELU, ACT_CLAMP01 = range(9)
CONV_OUT_F32, CONV_FORCE_SIMT, CONV_A_IM2COL, CONV_IN_OVERLAP, CONV_POOL3S2, CONV_F32_SPLIT, CONV_SE_GATE = 1, 2, 4, 8, 16, 32, 64

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_DEVICE = 0, -1, -2, -3, -4


class PcvError(RuntimeError):
    """A non-zero status from the C ABI (message from pcv_last_error)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"pcv_b200 error {code}: {msg}")
        self.code = code


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "N", "H", "W", "Cin", "Cout", "kh", "kw", "stride", "pad", "dil", "groups", "act",
        "in_pitch", "out_pitch", "res_pitch", "flags", "in_row_pitch")] + [("act_param", C.c_float)]


_P = C.c_void_p
_I = C.c_int
_Z = C.c_size_t

# name -> (restype, argtypes); must list every PCV_API symbol of include/pcv_b200.h (tests check this).
SIGNATURES = {
    "pcv_last_error": (C.c_char_p, []),
    "pcv_version": (_I, []),
    "pcv_device_info": (_I, [_I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_Z)]),
    "pcv_launch_count": (C.c_int64, []),
    "pcv_conv_packed_bytes": (_I, [C.POINTER(ConvDesc), _I, C.POINTER(_Z), C.POINTER(_Z)]),
    "pcv_pack_conv_weights": (_I, [C.POINTER(ConvDesc), _I, _P, _P, _P, _P, _P, _P, C.c_float, _P, _P, _P]),
    "pcv_conv2d_bias_act": (_I, [_P, C.POINTER(ConvDesc), _I, _P, _P, _P, _P, _P, _P]),
    "pcv_conv_se_gate_ok": (_I, [C.POINTER(ConvDesc), _I]),
    "pcv_conv_workspace_bytes": (_I, [C.POINTER(ConvDesc), _I, C.POINTER(_Z)]),
    "pcv_conv2d_bias_act_ws": (_I, [_P, C.POINTER(ConvDesc), _I, _P, _P, _P, _P, _P, _P, _P]),
    "pcv_bottleneck_tail_fusable": (_I, [C.POINTER(ConvDesc), C.POINTER(ConvDesc), _I]),
    "pcv_bottleneck_tail": (_I, [_P, C.POINTER(ConvDesc), C.POINTER(ConvDesc), _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "pcv_dw_pw_fusable": (_I, [C.POINTER(ConvDesc), C.POINTER(ConvDesc), _I]),
    "pcv_dw_pw_fused": (_I, [_P, C.POINTER(ConvDesc), C.POINTER(ConvDesc), _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "pcv_conv1x1_dual_ok": (_I, [C.POINTER(ConvDesc), C.POINTER(ConvDesc), _I]),
    "pcv_conv1x1_dual": (_I, [_P, C.POINTER(ConvDesc), C.POINTER(ConvDesc), _I, _P, _P, _P, _P, _P, _P]),
    "pcv_conv1x1_dual_se": (_I, [_P, C.POINTER(ConvDesc), C.POINTER(ConvDesc), _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "pcv_exp_dw_pw_fusable": (_I, [C.POINTER(ConvDesc), C.POINTER(ConvDesc), C.POINTER(ConvDesc), _I]),
    "pcv_exp_dw_pw_fused": (_I, [_P, C.POINTER(ConvDesc), C.POINTER(ConvDesc), C.POINTER(ConvDesc), _I, _P, _P, _P, _P, _P, _P,
                                _P, _P, _P, _P]),
    "pcv_zero_pad2d": (_I, [_P, _I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _P, _I, _P]),
    "pcv_maxpool2d": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P, _I, _P]),
    "pcv_global_avgpool": (_I, [_P, _I, _I, _I, _I, _P, _I, _P, _I, _P]),
    "pcv_adaptive_avgpool": (_I, [_P, _I, _I, _I, _I, _I, _P, _I, _I, _I, _P, _P]),
    "pcv_se_excite": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P, _I, _I, _P, _P]),
    "pcv_se_excite_ex": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _I, _I, _P, _P]),
    "pcv_se_scale_add_act": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _I, _P, _P]),
    "pcv_add_act": (_I, [_P, _I, _Z, _P, _P, _I, _P, _P]),
    "pcv_channel_affine_act": (_I, [_P, _I, _Z, _I, _P, _I, _P, _P, _P, _I, _P, _I, _P]),
    "pcv_nchw_f32_to_nhwc": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _I, _P]),
    "pcv_nchw_to_nhwc_ex": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P]),
    "pcv_nhwc_to_nchw_f32": (_I, [_P, _I, _I, _I, _I, _I, _P, _I, _P, _P]),
    "pcv_bilinear_upsample_ac": (_I, [_P, _I, _I, _I, _I, _I, _P, _I, _I, _I, _P, _I, _I, _P]),
    "pcv_stem_s2d_dims": (_I, [_I, _I, _I, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "pcv_stem_s2d_ingest": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "pcv_stem_s2d_ingest_ex": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "pcv_stem_s2d_weights": (_I, [_I, _I, _I, _P, _P, _P]),
    "pcv_stem_s2d_pool_ok": (_I, [_I, _I, _I, _I, _I]),
    "pcv_peer_buffer_bytes": (_I, [_I, _Z, C.POINTER(_Z)]),
    "pcv_peer_buffer_alloc": (_I, [_Z, C.POINTER(_P), _P]),
    "pcv_peer_buffer_open": (_I, [_P, C.POINTER(_P)]),
    "pcv_peer_buffer_close": (_I, [_P]),
    "pcv_peer_buffer_free": (_I, [_P]),
    "pcv_peer_allgather": (_I, [_P, _P, _Z, _I, _I, C.POINTER(_P), _P, _P]),
    "pcv_plan_create": (_I, [C.POINTER(_P)]),
    "pcv_plan_destroy": (_I, [_P]),
    "pcv_plan_num_ops": (_I, [_P]),
    "pcv_plan_num_launches": (_I, [_P]),
    "pcv_plan_run": (_I, [_P, _P]),
    "pcv_plan_graph_launch": (_I, [_P, _P]),
    "pcv_plan_profile": (_I, [_P, _P, C.POINTER(C.c_float), _I]),
    "pcv_plan_op_name": (C.c_char_p, [_P, _I]),
    "pcv_plan_op_cost": (_I, [_P, _I, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

_lib = None


def load() -> C.CDLL:
    """Load libpcv_b200.so (once).  Raises ImportError with build instructions when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C pytorchcv_b200/csrc`. pytorchcv_b200 has no CPU/PyTorch fallback path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        raise PcvError(status, load().pcv_last_error().decode("utf-8", "replace"))


def call(name: str, *args):
    """Call an int-status entry point and raise PcvError on failure."""
    check(getattr(load(), name)(*args))


def device_info(dev: int = 0):
    arch, sms, mem = _I(), _I(), _Z()
    call("pcv_device_info", dev, C.byref(arch), C.byref(sms), C.byref(mem))
    return arch.value, sms.value, mem.value


def launch_count() -> int:
    return int(load().pcv_launch_count())
