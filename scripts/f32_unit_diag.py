"""Per-unit diagnosis of the fp32 tier: every top-level block of a network is run standalone on the oracle's own input for
that block (tensor-core split route vs the oracle), so a wrong layer shows up without the network's error amplification."""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorchcv_b200 as P
from oracle import oracle_forward, seeded_init, seeded_input

name = sys.argv[1] if len(sys.argv) > 1 else "seresnext50_32x4d"
rb = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
net = seeded_init(P.get_model(name, pretrained=False).eval(), seed=0, randomize_bn=rb)
x = seeded_input((2, 3, 224, 224), seed=1234)
blocks = []
for nm, st in net.features.named_children():
    if nm.startswith("stage"):
        for un, u in st.named_children():
            blocks.append((f"{nm}.{un}", u))
    else:
        blocks.append((nm, st))
chain = len(sys.argv) > 3 and sys.argv[3] == "chain"   # feed each block the GPU path's own previous output
cur = x
gcur = x
for nm, blk in blocks:
    if isinstance(blk, (torch.nn.AvgPool2d, torch.nn.AdaptiveAvgPool2d)):
        break
    want = oracle_forward(blk, cur)
    got = P.accelerate(copy.deepcopy(blk).cuda(), dtype="fp32", graph=False)((gcur if chain else cur).cuda()).float().cpu()
    gcur = got
    rel = float((got - want).abs().max() / (want.abs().max() + 1e-30))
    print(f"{nm:18s} in {tuple(cur.shape)} rel {rel:.3e} max|ref| {float(want.abs().max()):.3e}", flush=True)
    cur = want
