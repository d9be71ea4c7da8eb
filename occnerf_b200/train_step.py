"""One training iteration of the reference's trainer (core/train/trainers/occnerf/trainer.py:223-253: Network.forward,
loss, backward, grad-norm clip, Adam, point_counter update) captured ONCE as a CUDA graph and replayed per step.

The step launches ~300 kernels (60 of ours + the library prologue + optimizer); issued from Python it is host-bound by
~2 ms on a B200 (measured: 13.6 ms of device work per 15.8 ms step).  All shapes on the path are static for a fixed ray
budget, so the whole iteration replays from one cudaGraphLaunch: inputs are copied into static device buffers from pinned
host memory, the loss comes back through a pinned scalar.

Constraints (checked or documented): fixed number of rays and samples; `iter_val`-dependent host decisions (Hann window,
non-rigid kick-in, pose-refinement kick-in) are baked in, so re-capture when the iteration crosses one of those
thresholds (`needs_recapture`); the optimizer must be capturable (Adam(fused=True, capturable=True)).
"""
from __future__ import annotations

import torch


class GraphedTrainStep:
    def __init__(self, net, optimizer, loss_fn, host_batch: dict, iter_val: int, params=None, max_norm: float | None = 1.0, warmup: int = 3,
                 grad_sync=None):
        """host_batch: dict of pinned host tensors with the keys of `Network.forward`'s per-frame inputs
        (rays_o, rays_d, near, far, dst_Rs, dst_Ts, cnl_gtfms, priors, posevec, bmin, bscale, bg) + whatever `loss_fn`
        needs (e.g. target).  loss_fn(out_dict, static_batch) -> scalar tensor.
        optimizer: occnerf_b200.optim.ClipAdam (clips inside its step; pass max_norm=None) or a capturable torch optimizer
        (then `max_norm` is applied with clip_grad_norm_ first).
        grad_sync(grads, hits): data-parallel hook between backward() and the optimizer.  A hook with `capturable = True`
        (distributed.SwitchReducer: the all-reduce is one of our kernels, the ranks meet inside it) is captured with everything else:
        ONE graph per step at every rank count.  Any other hook (NCCL) splits the iteration into TWO graphs -- forward + backward, and
        clip + optimizer + visibility update -- with the collectives launched eagerly in between.  (Capturing NCCL inside the step
        graph hung at 2 GPUs in round 1.)"""
        self.net, self.opt, self.loss_fn, self.iter_val, self.max_norm = net, optimizer, loss_fn, iter_val, max_norm
        self.grad_sync = grad_sync
        self.params = params if params is not None else [p for p in net.parameters() if p.requires_grad]
        dev = next(net.parameters()).device
        self.static = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in host_batch.items()}
        self.loss_dev = torch.zeros(1, device=dev)
        self.loss_host = torch.zeros(1).pin_memory()
        self.launches = 0
        self._load(host_batch)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                 # warm-up off the capture: builds caches, cuDNN plans, func attributes
            for _ in range(warmup):
                self._forward_backward()
                if self.grad_sync is not None:
                    self.grad_sync([p.grad for p in self.params], self._hits)
                self._optimize()
                self.opt.zero_grad(set_to_none=True)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.opt.zero_grad(set_to_none=True)
        from occnerf_b200 import _lib
        c0 = _lib.COUNTERS["launches"]
        self.graph = torch.cuda.CUDAGraph()
        if self.grad_sync is None or getattr(self.grad_sync, "capturable", False):
            with torch.cuda.graph(self.graph):
                self._forward_backward()
                if self.grad_sync is not None:
                    self.grad_sync([p.grad for p in self.params], self._hits)
                self._optimize()
                self.opt.zero_grad(set_to_none=True)
            self.graph_opt = None
        else:
            with torch.cuda.graph(self.graph):
                self._forward_backward()
            # the gradient tensors the backward graph writes (static addresses inside the graph's pool): reduced in place between
            # the two graphs, read by the optimizer graph -- so they are never released
            self.grads = [p.grad for p in self.params]
            self.graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt, pool=self.graph.pool()):
                self._optimize()
        self.launches = _lib.COUNTERS["launches"] - c0

    def _load(self, host_batch):
        for k, v in host_batch.items():
            self.static[k].copy_(v, non_blocking=True)

    def _forward_backward(self):
        d, net = self.static, self.net
        net.zero_bound_grads()
        out = net.forward((d["rays_o"], d["rays_d"]), d["dst_Rs"], d["dst_Ts"], d["cnl_gtfms"], d["priors"], dst_posevec=d["posevec"],
                          near=d["near"], far=d["far"], iter_val=self.iter_val, cnl_bbox_min_xyz=d["bmin"], cnl_bbox_scale_xyz=d["bscale"],
                          bgcolor=d["bg"])
        loss = self.loss_fn(out, d)
        loss.backward()
        net.attach_bound_grads()
        self._hits = out.get("hits")
        self.loss_dev.copy_(loss.detach().reshape(1))

    def _optimize(self):
        if self.max_norm is not None:
            torch.nn.utils.clip_grad_norm_(self.params, self.max_norm)
        self.opt.step()
        if self._hits is not None:
            self.net.apply_visibility(self._hits)

    def _iteration(self):
        """The whole iteration, eagerly (what the graph(s) replay)."""
        self._forward_backward()
        if self.grad_sync is not None:
            self.grad_sync([p.grad for p in self.params], self._hits)
        self._optimize()
        self.opt.zero_grad(set_to_none=True)

    def needs_recapture(self, iter_val: int) -> bool:
        cfg = self.net.cfg
        marks = (cfg.non_rigid_kick_in_iter, cfg.non_rigid_full_band_iter, getattr(self.net, "_pose_kick_in_iter", float("inf")))
        return any((self.iter_val < m) != (iter_val < m) for m in marks) or \
            (cfg.non_rigid_kick_in_iter <= iter_val < cfg.non_rigid_full_band_iter and iter_val != self.iter_val)

    def step(self, host_batch: dict) -> torch.Tensor:
        """H2D of the frame -> graph launch(es) -> D2H of the loss (returned as a pinned host tensor; valid after a sync)."""
        self._load(host_batch)
        self.graph.replay()
        if self.graph_opt is not None:
            self.grad_sync(self.grads, self._hits)
            self.graph_opt.replay()
        self.loss_host.copy_(self.loss_dev, non_blocking=True)
        return self.loss_host
