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
    """Sum-all-reduce of a fixed list of gradient tensors with the fewest bytes and launches:
      * tensors below `bucket_bytes` travel through ONE persistent flat buffer (a multi-tensor copy in, one collective, a
        multi-tensor copy out -- no torch.cat, no per-tensor copy-back, no division pass: the caller folds 1 / world_size into
        the loss); the 0/1 visibility votes ride in the same buffer (their sum > 0 <=> some rank voted, network.py:517);
      * the large ones (hash table, decoder weights) are reduced in place;
      * `active` maps a tensor (by position in the list) to the index expression of the only entries that can be non-zero on
        ANY rank; only that compact block is exchanged.  Used for the decoder's first ConvTranspose3d(1024, 512, 4, 2, 1): its
        input is 1x1x1, so 56 of its 64 taps never receive a gradient (117 of the 134 MB of that tensor are structural zeros);
      * all collectives of a step are issued as ONE NCCL group (a single launch) and waited for together."""

    def __init__(self, bucket_bytes: int = 4 << 20, active: dict | None = None):
        self.bucket_bytes = bucket_bytes
        self.active = dict(active or {})
        self.flat, self.views, self.key = None, None, None

    def __call__(self, grads, hits=None):
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        compact = {}
        items = []
        for i, g in enumerate(grads):
            if g is None:
                continue
            if i in self.active:
                c = g[self.active[i]].contiguous()
                compact[i] = c
                items.append(c)
            else:
                items.append(g)
        if hits is not None:
            items.append(hits)
        small = [g for g in items if g.numel() * g.element_size() < self.bucket_bytes]
        large = [g for g in items if g.numel() * g.element_size() >= self.bucket_bytes]
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
        tensors = ([self.flat] if small else []) + large
        if dist.get_backend() == "nccl" and hasattr(dist, "_coalescing_manager") and len(tensors) > 1:
            with dist._coalescing_manager(device=tensors[0].device, async_ops=True) as cm:     # one NCCL group = one launch
                for t in tensors:
                    dist.all_reduce(t, op=dist.ReduceOp.SUM)
            cm.wait()
        else:                                                  # (gloo in the CPU tests: no coalescing)
            for w in [dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True) for t in tensors]:
                w.wait()
        if small:
            torch._foreach_copy_(small, self.views)
        for i, c in compact.items():
            grads[i][self.active[i]] = c
        if hits is not None:
            hits.clamp_(max=1.0)                               # votes are 0/1: any rank's vote counts once


import os as _os
_SKIP_KERNEL = _os.environ.get("OCCNERF_SKIP_ALLREDUCE", "0") != "0"


class SwitchReducer:
    """GradReducer's job as ONE hand-written kernel over NVSwitch peer memory (csrc/collective.cu, occnerf_allreduce_sum_f32)
    instead of NCCL calls: a two-shot sum in which the switch itself adds the ranks' copies (multimem.ld_reduce / multimem.st on the
    NVLS multicast mapping; peer loads and stores where no multicast mapping exists).

    All gradients of a step live in one flat fp32 buffer in symmetric memory (torch.distributed._symmetric_memory does the
    allocation and the handle exchange -- plumbing; the data path is the kernel):
        [ table | bucket | reserved ]
      * `table_view` is bound as `Network.emb_grad_out`: occnerf_hashgrid_backward scatters the 59 MiB table gradient straight into
        it (no copy in, no copy out; the owner zeroes it at the start of a step);
      * `reserve(shape)` hands out further in-place destinations (the decoder's weight gradients: occnerf_deconv3d_backward writes
        them there, prologue._DecoderFn.grad_out); any gradient that already lives inside the buffer is left where it is;
      * every other gradient (MLP, point_dist, weight volume / decoder weights, compacted where structurally zero) and the 0/1
        visibility votes are copied into the bucket by one multi-tensor copy and back by another.
    The ranks synchronise inside the kernel, not on the host and not through a communicator stream, so the launch sits in the
    compute stream like any other kernel and the whole training step -- collective included -- replays from ONE CUDA graph at
    every rank count."""

    capturable = True

    def __init__(self, table_numel: int, bucket_numel: int, device, group=None, blocks: int = 16, active: dict | None = None,
                 reserve_numel: int = 0):
        import torch.distributed._symmetric_memory as symm
        from occnerf_b200 import _lib
        _lib.load()
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > 8:
            raise RuntimeError("SwitchReducer: at most 8 ranks (one NVSwitch domain)")
        self.active = dict(active or {})
        blocks = int(_os.environ.get("OCCNERF_AR_BLOCKS", blocks))
        self.blocks = blocks
        self.table_numel = (table_numel + 3) // 4 * 4
        self.bucket_end = self.table_numel + (bucket_numel + 3) // 4 * 4         # [table_numel, bucket_end): the copy-in bucket
        self.reserve_pos = self.bucket_end                                        # [bucket_end, capacity): in-place destinations
        self.capacity = self.bucket_end + (reserve_numel + 3) // 4 * 4
        self.flat = symm.empty(self.capacity, dtype=torch.float32, device=device)
        self.flat.zero_()
        self.hdl = symm.rendezvous(self.flat, self.group)
        self.pad = symm.empty(blocks * 8, dtype=torch.int32, device=device)
        self.pad.zero_()
        self.hdl_pad = symm.rendezvous(self.pad, self.group)
        self.epochs = torch.zeros(blocks, dtype=torch.int32, device=device)
        self.multicast = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        bufs, pads = list(self.hdl.buffer_ptrs), list(self.hdl_pad.buffer_ptrs)
        if bufs[self.rank] != self.flat.data_ptr() or pads[self.rank] != self.pad.data_ptr():
            raise RuntimeError("SwitchReducer: symmetric-memory handle does not start at the tensor (unexpected allocator offset)")
        import ctypes as C
        self._bufs = (C.c_void_p * self.world)(*bufs)
        self._pads = (C.c_void_p * self.world)(*pads)
        self.table_view = self.flat[:table_numel]
        self.bucket = self.flat[self.table_numel:self.bucket_end]
        self.views, self.key = None, None
        torch.cuda.synchronize(device)
        dist.barrier(self.group)                  # every rank's pad and buffer are zeroed before anybody signals into them
        self.kind = "nvls-multimem" if self.multicast else "p2p"

    def bind_table(self, net):
        """Routes the embeddings' gradient of `net` into the symmetric buffer (zero it with `zero_table()` before each backward)."""
        net.emb_grad_out = self.table_view

    def zero_table(self):
        self.table_view.zero_()

    def reserve(self, shape):
        """A gradient destination of `shape` inside the all-reduce buffer (16-byte aligned); whoever produces that gradient writes it
        here and it is reduced in place."""
        n = 1
        for d in shape:
            n *= int(d)
        if self.reserve_pos + n > self.capacity:
            raise RuntimeError("SwitchReducer.reserve: the reserved region is full")
        view = self.flat[self.reserve_pos:self.reserve_pos + n].view(*shape)
        self.reserve_pos += (n + 3) // 4 * 4
        return view

    def _inside(self, t):
        """0: not in the buffer; 1: the table gradient; 2: a reserved in-place destination."""
        off = (t.data_ptr() - self.flat.data_ptr()) // 4
        if 0 <= off < self.table_numel:
            return 1
        return 2 if self.bucket_end <= off < self.capacity else 0

    def __call__(self, grads, hits=None):
        import ctypes as C
        from occnerf_b200._lib import call, stream
        compact, items, reserved_used = {}, [], False
        for i, g in enumerate(grads):
            if g is None:
                continue
            where = self._inside(g)
            if where:                                           # already in place (table gradient, reserved destinations)
                reserved_used |= where == 2
                continue
            if i in self.active:
                c = g[self.active[i]].contiguous()
                compact[i] = c
                items.append(c)
            else:
                items.append(g)
        if hits is not None:
            items.append(hits)
        key = tuple((tuple(g.shape), g.dtype) for g in items) + (reserved_used,)
        if key != self.key:
            n = sum(g.numel() for g in items)
            if n > self.bucket.numel():
                raise RuntimeError(f"SwitchReducer: bucket of {self.bucket.numel()} floats is too small for {n}")
            self.views, off = [], 0
            for g in items:
                self.views.append(self.bucket[off:off + g.numel()].view_as(g))
                off += g.numel()
            # one contiguous prefix is reduced: up to the end of the used part of the bucket, or -- when gradients of this call live in
            # the reserved region -- up to the end of what has been reserved
            self.used = self.reserve_pos if reserved_used else self.table_numel + (off + 3) // 4 * 4
            self.key = key
        if items:
            torch._foreach_copy_(self.views, items)
        if not _SKIP_KERNEL:          # (OCCNERF_SKIP_ALLREDUCE=1: timing experiment -- everything but the collective itself)
            call("occnerf_allreduce_sum_f32", C.cast(self._bufs, C.c_void_p), C.cast(self._pads, C.c_void_p), self.multicast or None,
                 self.used, self.rank, self.world, self.blocks, self.epochs.data_ptr(), stream())
        if items:
            torch._foreach_copy_(items, self.views)
        for i, c in compact.items():
            grads[i][self.active[i]] = c
        if hits is not None:
            hits.clamp_(max=1.0)                               # votes are 0/1: any rank's vote counts once


def structural_zero_slices(params):
    """{position: index expression} for GradReducer(active=...): parameters whose gradient is structurally zero outside a block.
    A ConvTranspose3d(k=4, s=2, p=1) weight [Cin, Cout, 4, 4, 4] fed by a 1x1x1 input only ever uses taps 1..2 per axis."""
    out = {}
    for i, p in enumerate(params):
        if p.dim() == 5 and tuple(p.shape[2:]) == (4, 4, 4) and p.shape[0] == 1024:
            out[i] = (slice(None), slice(None), slice(1, 3), slice(1, 3), slice(1, 3))
    return out


def allreduce_visibility(hits: torch.Tensor) -> torch.Tensor:
    """A point is voted for if any rank's rays voted for it (duplicates collapse upstream too)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(hits, op=dist.ReduceOp.MAX)
    return hits
