"""Data-parallel plumbing for the ray path: one process per GPU, rays sharded by contiguous index range, no
data-path collective.  Training adds one gradient all-reduce between `backward()` and the
clip/optimizer step (trainer.py:246-248) and a max-reduce of the visibility votes (network.py:517 is state, not a
gradient).  Works with any torch.distributed backend (nccl on the B200 box, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int, granule: int = 1):
    """Contiguous [begin, end) of `n_items` for `rank`, in multiples of `granule` (e.g. 1024 = one 32x32 patch so that
    whole patches stay on one rank, trainer.py:31-41).  Every item is covered exactly once; the tail goes to the last ranks."""
    units = (n_items + granule - 1) // granule
    base, rem = divmod(units, world)
    u0 = rank * base + min(rank, rem)
    u1 = u0 + base + (1 if rank < rem else 0)
    return min(u0 * granule, n_items), min(u1 * granule, n_items)


def shard_rays(rays: torch.Tensor, rank: int, world: int, granule: int = 1) -> torch.Tensor:
    b, e = shard_range(rays.shape[0], rank, world, granule)
    return rays[b:e]


def gather_rays(local: torch.Tensor, n_total: int, rank: int, world: int, granule: int = 1) -> torch.Tensor:
    """Host-side assembly of per-rank outputs (image assembly is outside the timed path)."""
    if world == 1:
        return local
    sizes = [shard_range(n_total, r, world, granule) for r in range(world)]
    out = [torch.empty((e - b,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device) for b, e in sizes]
    dist.all_gather(out, local.contiguous())
    return torch.cat(out, 0)


def allreduce_gradients(params, extra=(), average: bool = True, bucket_bytes: int = 64 << 20):
    """Sum (or average) `.grad` of `params` and the tensors in `extra` over all ranks, small tensors coalesced into
    buckets so that the 20 MLP tensors travel as one message while the 59 MiB hash-table gradient goes alone."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    world = dist.get_world_size()
    tensors = [p.grad for p in params if p.grad is not None] + [t for t in extra if t is not None]
    small, work = [], []
    for t in tensors:
        if t.numel() * t.element_size() >= bucket_bytes // 4:
            work.append((dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True), [t], None))
        else:
            small.append(t)
    if small:
        flat = torch.cat([t.reshape(-1) for t in small])
        work.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True), small, flat))
    for w, ts, flat in work:
        w.wait()
        if flat is not None:
            off = 0
            for t in ts:
                t.copy_(flat[off:off + t.numel()].view_as(t))
                off += t.numel()
        if average:
            for t in ts:
                t.div_(world)


class GradReducer:
    """Sum-all-reduce of a fixed list of gradient tensors with the fewest launches: tensors below `bucket_bytes` travel through ONE
    persistent flat buffer (a multi-tensor copy in, one collective, a multi-tensor copy out -- no torch.cat, no per-tensor
    copy-back, no division pass: the caller folds 1 / world_size into the loss), the large ones (hash table, decoder weights)
    are reduced in place.  All collectives are issued asynchronously and waited for together, so NCCL pipelines them."""

    def __init__(self, bucket_bytes: int = 4 << 20):
        self.bucket_bytes = bucket_bytes
        self.flat, self.views, self.key = None, None, None

    def __call__(self, grads):
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        grads = [g for g in grads if g is not None]
        small = [g for g in grads if g.numel() * g.element_size() < self.bucket_bytes]
        large = [g for g in grads if g.numel() * g.element_size() >= self.bucket_bytes]
        work = []
        if small:
            key = tuple((tuple(g.shape), g.dtype, g.device) for g in small)
            if key != self.key:
                self.flat = torch.empty(sum(g.numel() for g in small), dtype=small[0].dtype, device=small[0].device)
                self.views, off = [], 0
                for g in small:
                    self.views.append(self.flat[off:off + g.numel()].view_as(g))
                    off += g.numel()
                self.key = key
            torch._foreach_copy_(self.views, small)
            work.append(dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=True))
        for g in large:
            work.append(dist.all_reduce(g, op=dist.ReduceOp.SUM, async_op=True))
        for w in work:
            w.wait()
        if small:
            torch._foreach_copy_(small, self.views)


def allreduce_visibility(hits: torch.Tensor) -> torch.Tensor:
    """A point is voted for if any rank's rays voted for it (duplicates collapse upstream too)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(hits, op=dist.ReduceOp.MAX)
    return hits
