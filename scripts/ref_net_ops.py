"""Op-name histogram of a network lowered from the reference's own modules (baseline/_ref or /root/reference): which kernels serve it."""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
for path in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
    if os.path.isdir(os.path.join(path, "pytorchcv")):
        sys.path.insert(0, path)
        break
import pytorchcv_b200 as P  # noqa: E402
from pytorchcv.model_provider import get_model  # noqa: E402

for name in sys.argv[1:]:
    net = get_model(name, pretrained=False).eval().cuda()
    x = torch.randn(8, 3, 224, 224, device="cuda")
    fast = P.accelerate(net, dtype="bf16")
    fast(x)
    rows = fast.compiled(x).profile()
    hist = collections.Counter(r[0].split(" ")[0] for r in rows)
    print(name, len(rows), "ops", sum(r[1] for r in rows), "ms:", dict(hist))
