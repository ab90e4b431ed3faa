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


class PeerExchange:
    """The logits all-gather as ONE device-initiated kernel over NVLink peer memory (include/pcv_b200.h, "multi-GPU
    exchange"): every rank owns an exchange buffer exported through CUDA IPC; torch.distributed only carries the 64-byte
    handles once at set-up.  Raises if any rank cannot map every peer (callers fall back to NCCL on ALL ranks)."""

    MAX_BYTES = 16 << 20   # per rank: a 64-CTA kernel sized for logits; large segmentation maps go through NCCL

    def __init__(self, rank: int, world: int, bytes_per_rank: int, device: torch.device, group=None):
        import ctypes as C
        from . import _lib
        self.rank, self.world, self.bytes, self.device = rank, world, bytes_per_rank, device
        self._lib, self._C = _lib, C
        self._own, self._opened = None, []
        ok, err = 1, ""
        ptrs = [None] * world
        try:
            if bytes_per_rank % 16 or bytes_per_rank > self.MAX_BYTES or world > 16:
                raise ValueError(f"payload of {bytes_per_rank} B per rank is outside the peer kernel's domain")
            total = C.c_size_t()
            _lib.call("pcv_peer_buffer_bytes", world, bytes_per_rank, C.byref(total))
            own, handle = C.c_void_p(), C.create_string_buffer(64)
            with torch.cuda.device(device):
                _lib.call("pcv_peer_buffer_alloc", total.value, C.byref(own), handle)
            self._own = own
            handles = [None] * world
            dist.all_gather_object(handles, handle.raw, group=group)
            ptrs[rank] = own.value
            with torch.cuda.device(device):
                for r in range(world):
                    if r != rank:
                        q = C.c_void_p()
                        _lib.call("pcv_peer_buffer_open", C.create_string_buffer(handles[r], 64), C.byref(q))
                        self._opened.append(q)
                        ptrs[r] = q.value
        except Exception as e:  # noqa: BLE001
            ok, err = 0, repr(e)
            if self._own is None:   # keep the collective call pattern identical on every rank
                try:
                    dist.all_gather_object([None] * world, b"", group=group)
                except Exception:  # noqa: BLE001
                    pass
        flag = torch.tensor([ok], device=device if dist.get_backend(group) == "nccl" else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            self.close()
            raise RuntimeError(f"peer exchange unavailable on at least one rank ({err or 'another rank failed'})")
        self._ptrs = (C.c_void_p * world)(*ptrs)

    def gather(self, local: torch.Tensor) -> torch.Tensor:
        local = local.contiguous()
        if local.numel() * local.element_size() != self.bytes:
            raise ValueError("peer exchange was sized for a different shard")
        out = local.new_empty((self.world * local.shape[0],) + tuple(local.shape[1:]))
        with torch.cuda.device(self.device):
            self._lib.call("pcv_peer_allgather", None, local.data_ptr(), self.bytes, self.rank, self.world, self._ptrs,
                           out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
        return out

    def close(self) -> None:
        lib = self._lib
        try:
            with torch.cuda.device(self.device):
                torch.cuda.synchronize(self.device)
                for q in self._opened:
                    lib.load().pcv_peer_buffer_close(q)
                if self._own is not None:
                    lib.load().pcv_peer_buffer_free(self._own)
        except Exception:  # noqa: BLE001
            pass
        self._opened, self._own = [], None

    def __del__(self):
        self.close()


class ShardedInference:
    """Run `net` (an accelerated module) on this rank's shard of a global batch and gather the logits.

    exchange: "peer" = the device-initiated NVLink kernel (equal shards of a single [n, C] CUDA tensor), "nccl" = one
    torch.distributed all-gather, "auto" = peer when every rank can map every peer's buffer, else NCCL (decided jointly)."""

    def __init__(self, net, rank: int, world: int, group=None, exchange: str = "auto"):
        if exchange not in ("auto", "peer", "nccl"):
            raise ValueError(f"unknown exchange {exchange!r}")
        self.net, self.rank, self.world, self.group = net, rank, world, group
        self.exchange, self.exchange_used, self._peer = exchange, None, None

    def local_slice(self, global_batch: int) -> slice:
        lo, hi = shard_bounds(global_batch, self.world, self.rank)
        return slice(lo, hi)

    def _peer_for(self, y) -> PeerExchange | None:
        """The peer exchange for this output, set up on first use; None when the output / group is outside its domain."""
        if self._peer is not None:
            return self._peer
        if self.exchange == "nccl" or (self.exchange_used or "").startswith("nccl"):
            self.exchange_used = "nccl: one all_gather_into_tensor per output"
            return None
        eligible = isinstance(y, torch.Tensor) and y.is_cuda and y.dim() == 2
        if eligible:
            try:
                self._peer = PeerExchange(self.rank, self.world, y.numel() * y.element_size(), y.device, self.group)
                self.exchange_used = "peer: one device-initiated NVLink kernel (push / flag / wait / drain)"
                return self._peer
            except RuntimeError:
                if self.exchange == "peer":
                    raise
        elif self.exchange == "peer":
            raise RuntimeError("exchange='peer' needs a single [n, C] CUDA tensor per rank")
        self.exchange_used = "nccl: one all_gather_into_tensor per output"
        return None

    def __call__(self, x_local: torch.Tensor, global_batch: int | None = None):
        y = self.net(x_local)
        if self.world == 1:
            return y
        sizes = None
        if global_batch is not None:
            sizes = [hi - lo for lo, hi in (shard_bounds(global_batch, self.world, r) for r in range(self.world))]
        if sizes is None or len(set(sizes)) == 1:
            peer = self._peer_for(y)
            if peer is not None:
                return peer.gather(y)
        elif self.exchange_used is None:
            self.exchange_used = "nccl: one all_gather_into_tensor per output"
        return gather_logits(y, sizes, self.group)
