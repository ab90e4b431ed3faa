import torch, pytorchcv_b200 as P
from pytorchcv_b200 import nets as M, blocks as BK
unit = M.LinearBottleneck(24, 24, 1, expansion=True, remove_exp_conv=True, activation=BK.lambda_relu6()).eval().cuda()
x = torch.randn(256, 24, 56, 56, device="cuda")
fast = P.accelerate(unit, dtype="fp16", graph=False)
fast(x); torch.cuda.synchronize()
