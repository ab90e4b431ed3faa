"""Batch-sharded multi-GPU inference: one process per GPU, independent replicas, one all-gather of logits.

Eval-mode images are independent (BatchNorm uses running statistics), so the path shards by batch with no data-path
collective; the only exchange is the final all-gather of the [N/G, classes] logits.  `torch.distributed`
(NCCL over NVLink on GPUs, gloo in the CPU tests) is plumbing only.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Rank r owns images [lo, hi): contiguous, sizes differ by at most one, earlier ranks take the remainder."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's env (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def gather_logits(local, sizes: list[int] | None = None, group=None):
    """All-gather per-rank outputs [n_r, ...] into [sum n_r, ...] on every rank (rank order == image order).

    Equal shards use one all_gather_into_tensor (a single NCCL all-gather); ragged shards pad to the largest.
    Tuples / lists (DeepLabv3 / FCN / PSPNet with aux=True, ResNetD with multi_output) are gathered element-wise."""
    if isinstance(local, (tuple, list)):
        return type(local)(gather_logits(t, sizes, group) for t in local)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    local = local.contiguous()
    if sizes is None:
        sizes = [local.shape[0]] * world
    if len(set(sizes)) == 1:
        out = local.new_empty((world * local.shape[0],) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    width = max(sizes)
    padded = local.new_zeros((width,) + tuple(local.shape[1:]))
    padded[: local.shape[0]] = local
    out = local.new_empty((world * width,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * width: r * width + sizes[r]] for r in range(world)], dim=0)


class ShardedInference:
    """Run `net` (an accelerated module) on this rank's shard of a global batch and gather the logits."""

    def __init__(self, net, rank: int, world: int, group=None):
        self.net, self.rank, self.world, self.group = net, rank, world, group

    def local_slice(self, global_batch: int) -> slice:
        lo, hi = shard_bounds(global_batch, self.world, self.rank)
        return slice(lo, hi)

    def __call__(self, x_local: torch.Tensor, global_batch: int | None = None) -> torch.Tensor:
        y = self.net(x_local)
        if self.world == 1:
            return y
        sizes = None
        if global_batch is not None:
            sizes = [hi - lo for lo, hi in (shard_bounds(global_batch, self.world, r) for r in range(self.world))]
        return gather_logits(y, sizes, self.group)
