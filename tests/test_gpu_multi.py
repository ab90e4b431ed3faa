"""Two-GPU test of the device-initiated logits all-gather (pcv_peer_allgather) and of sharded inference end to end.
Skipped on single-GPU boxes; the host logic has its gloo twin in tests/test_dist.py."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import pytorchcv_b200 as P
    from pytorchcv_b200 import parallel
    from oracle import seeded_init, seeded_input
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, local = parallel.init_from_env("nccl")
    dev = torch.device("cuda", local)
    ok = True
    # raw exchange: 12 steps of changing payloads, both slab parities, results in rank order
    step = {"i": 0}

    def fake(x):
        return (torch.arange(16 * 1000, device=dev, dtype=torch.float32).view(16, 1000) + 1e5 * r + 7.0 * step["i"])
    runner = parallel.ShardedInference(fake, r, w, exchange="peer")
    for i in range(12):
        step["i"] = i
        out = runner(None)
        want = torch.cat([torch.arange(16 * 1000, device=dev, dtype=torch.float32).view(16, 1000) + 1e5 * rr + 7.0 * i
                          for rr in range(w)])
        ok = ok and out.shape == (16 * w, 1000) and torch.equal(out, want)
    ok = ok and runner.exchange_used.startswith("peer")
    # whole network: batch 8 sharded over two replicas == the same 8 images on one replica
    net = seeded_init(P.get_model("resnet18", pretrained=False).eval(), seed=0).to(dev)
    fast = P.accelerate(net, dtype="bf16")
    x = seeded_input((8, 3, 224, 224), seed=5).to(dev)
    whole = fast(x)
    for exchange in ("peer", "nccl"):
        sharded = parallel.ShardedInference(fast, r, w, exchange=exchange)
        got = sharded(x[sharded.local_slice(8)], global_batch=8)
        # batch 4 and batch 8 plans may tile differently: compare within the tier's rounding, argmax exactly
        rel = float((got - whole).abs().max() / whole.abs().max())
        ok = ok and got.shape == whole.shape and rel <= 1e-2 and torch.equal(got.argmax(1), whole.argmax(1))
    torch.cuda.synchronize(dev)
    q.put((rank, bool(ok)))
    dist.barrier(device_ids=[local])
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_peer_allgather_and_sharded_inference_two_gpus():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(ok for _, ok in results), results
