"""Loss epilogue of the reference's training iteration on the device: `clip_grad_norm_(network.parameters(), 1.0)` followed by
`torch.optim.Adam.step()` (core/train/trainers/occnerf/trainer.py:248-249, optimizers/occnerf/optimizer.py:12-43) as ONE C call
(occnerf_clip_adam_step, csrc/optim.cu: two multi-tensor kernels, 32 B per parameter, no host sync, graph-capturable).

The class keeps torch.optim.Adam's surface where the reference touches it: `param_groups` (a list of dicts with "params", "lr",
"name" -- what lr_updaters/exp_decay.py rewrites every iteration), `step()`, `zero_grad()`, and `state_dict()` /
`load_state_dict()` in torch's own format, so optimizer state written by the reference's trainer (trainer.py:398-430) loads here
and vice versa."""
from __future__ import annotations

import ctypes as C

import torch

from occnerf_b200._lib import call, stream

f32 = torch.float32


class ClipAdam:
    def __init__(self, params, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, max_norm=1.0):
        groups = list(params)
        if groups and not isinstance(groups[0], dict):
            groups = [{"params": groups}]
        self.param_groups = []
        for g in groups:
            g = dict(g)
            g["params"] = list(g["params"])
            g.setdefault("lr", lr)
            g.setdefault("betas", betas)
            g.setdefault("eps", eps)
            self.param_groups.append(g)
        self.betas, self.eps, self.max_norm = betas, eps, max_norm
        self.state = {}
        self._state2 = None          # device double: squared gradient norm of the last step

    def _params(self):
        return [(p, g["lr"]) for g in self.param_groups for p in g["params"]]

    def _init(self, dev):
        if self._state2 is None:
            self._state2 = torch.zeros(1, device=dev, dtype=torch.float64)

    def zero_grad(self, set_to_none: bool = True):
        for p, _ in self._params():
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    @torch.no_grad()
    def step(self):
        todo = [(p, lr) for p, lr in self._params() if p.grad is not None]
        if not todo:
            return
        dev = todo[0][0].device
        self._init(dev)
        n = len(todo)
        P, G, M, V, T = (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_void_p * n)()
        numel, lrs = (C.c_long * n)(), (C.c_float * n)()
        keep = []
        for i, (p, lr) in enumerate(todo):
            if not (p.is_cuda and p.dtype == f32 and p.is_contiguous()):
                raise RuntimeError("ClipAdam: parameters must be contiguous fp32 CUDA tensors (there is no CPU path)")
            g = p.grad
            if g.dtype != f32 or not g.is_contiguous():
                g = g.contiguous().float()
                keep.append(g)
            st = self.state.get(p)
            if st is None:
                st = self.state[p] = {"step": torch.zeros((), device=dev, dtype=f32),
                                      "exp_avg": torch.zeros_like(p, memory_format=torch.contiguous_format),
                                      "exp_avg_sq": torch.zeros_like(p, memory_format=torch.contiguous_format)}
            P[i], G[i], M[i], V[i], T[i] = p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), st["step"].data_ptr()
            numel[i], lrs[i] = p.numel(), float(lr)
        call("occnerf_clip_adam_step", C.cast(P, C.c_void_p), C.cast(G, C.c_void_p), C.cast(M, C.c_void_p), C.cast(V, C.c_void_p),
             C.cast(T, C.c_void_p), C.cast(numel, C.c_void_p), C.cast(lrs, C.c_void_p), n, float(self.betas[0]), float(self.betas[1]), float(self.eps),
             float(self.max_norm if self.max_norm is not None else 0.0), self._state2.data_ptr(), stream())

    def grad_norm(self) -> torch.Tensor:
        """L2 norm of all gradients as seen by the last step() (device scalar; what clip_grad_norm_ returns)."""
        return self._state2[0].sqrt().float()

    # -- torch.optim.Adam checkpoint format
    def state_dict(self):
        idx, packed_groups, state = 0, [], {}
        for g in self.param_groups:
            ids = []
            for p in g["params"]:
                if p in self.state:
                    state[idx] = {"step": self.state[p]["step"].detach().cpu().clone(), "exp_avg": self.state[p]["exp_avg"],
                                  "exp_avg_sq": self.state[p]["exp_avg_sq"]}
                ids.append(idx)
                idx += 1
            pg = {k: v for k, v in g.items() if k != "params"}
            pg["params"] = ids
            packed_groups.append(pg)
        return {"state": state, "param_groups": packed_groups}

    def load_state_dict(self, sd):
        flat = [p for g in self.param_groups for p in g["params"]]
        for g, sg in zip(self.param_groups, sd["param_groups"]):
            for k, v in sg.items():
                if k != "params":
                    g[k] = v
        for i, st in sd["state"].items():
            p = flat[int(i)]
            self.state[p] = {"step": torch.as_tensor(float(st["step"]), dtype=f32).to(p.device),
                             "exp_avg": st["exp_avg"].to(p.device, f32).contiguous().clone(),
                             "exp_avg_sq": st["exp_avg_sq"].to(p.device, f32).contiguous().clone()}
        if flat:
            self._init(flat[0].device)
