"""HBM bandwidth probes on the GPU box: write-only (fill), read-only (sum), copy; sizes well above the 126 MB L2."""
import torch
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
ms = t(lambda: a.zero_()); print(f"fill  : {2*n/ms/1e6:8.1f} GB/s ({ms:.3f} ms)")
ms = t(lambda: a.fill_(1.5)); print(f"fill2 : {2*n/ms/1e6:8.1f} GB/s")
ms = t(lambda: b.copy_(a)); print(f"copy  : {4*n/ms/1e6:8.1f} GB/s total")
ai = a.view(torch.int32)
ms = t(lambda: ai.sum()); print(f"read  : {2*n/ms/1e6:8.1f} GB/s")
c = torch.empty(n // 4, dtype=torch.bfloat16, device="cuda")
ms = t(lambda: torch.add(a[: n // 4], b[: n // 4], out=c)); print(f"2r1w  : {3*(n//4)*2/ms/1e6:8.1f} GB/s total")
