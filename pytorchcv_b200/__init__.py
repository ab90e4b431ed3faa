"""pytorchcv_b200 — a B200-native (sm_100a) eval-mode convolution path for pytorchcv models.

    from pytorchcv_b200 import get_model, accelerate
    net = get_model("resnet50", pretrained=False).eval().cuda()   # mirror modules: forward runs the B200 plan
    y = net(x)                                                     # x: [N,3,224,224] fp32 CUDA tensor
    fast = accelerate(reference_net.eval().cuda())                 # or wrap a real `pytorchcv` module tree

Everything computes through libpcv_b200.so (hand-written CUDA behind the C ABI of include/pcv_b200.h); there is
no torch-op or CPU fallback.
"""
from . import _lib
from .model_provider import get_model, supported_models
from .plan import (Accelerated, CompiledModule, accelerate, invalidate, run_module, set_default_graph,
                   set_default_precision)

__all__ = ["get_model", "supported_models", "accelerate", "invalidate", "run_module", "set_default_precision", "set_default_graph",
           "CompiledModule", "Accelerated", "_lib"]
__version__ = "0.1.0"
