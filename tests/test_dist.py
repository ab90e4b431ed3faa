"""world_size-2 gloo test (CPU) of the multi-GPU host logic: batch sharding and the logits all-gather."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, global_batch, q):
    sys.path.insert(0, ROOT)
    from pytorchcv_b200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = parallel.init_from_env("gloo")
    full = torch.arange(global_batch * 5, dtype=torch.float32).view(global_batch, 5)
    runner = parallel.ShardedInference(lambda t: t * 2.0, r, w)       # stand-in replica: logits = 2 * x
    out = runner(full[runner.local_slice(global_batch)], global_batch=global_batch)
    ok = torch.equal(out, full * 2.0)
    # tuple outputs (DeepLabv3 / FCN / PSPNet with aux=True): gathered element-wise, structure preserved
    seg = parallel.ShardedInference(lambda t: (t + 1.0, [t - 1.0]), r, w, exchange="nccl")
    main, (aux,) = seg(full[runner.local_slice(global_batch)], global_batch=global_batch)
    ok = ok and torch.equal(main, full + 1.0) and torch.equal(aux, full - 1.0)
    ok = ok and runner.exchange_used.startswith("nccl")          # CPU tensors never take the peer-memory kernel
    q.put((rank, ok, tuple(out.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("global_batch", [8, 7])
def test_sharded_inference_gathers_logits_in_image_order(global_batch):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, global_batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, shape in results:
        assert ok and shape == (global_batch, 5), (rank, shape)


def test_shard_bounds_partition_the_batch():
    from pytorchcv_b200.parallel import shard_bounds
    for n in (1, 7, 8, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)
