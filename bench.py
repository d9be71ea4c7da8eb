#!/usr/bin/env python
"""Benchmark of the OccNeRF per-ray rendering path (BASELINE.json metric: rays/s at 128 samples/ray, fwd+bwd).

    python bench.py --gpus N --steps K --warmup W [--engine tf32|tc3|fp32|tc1] [--impl reference]

Workload (BASELINE.json configs[1]): one ZJU-Mocap-387-shaped training step = 6 patches of 32x32 rays
(6144 rays, 786 432 samples), synthetic SMPL-like subject, random-init network, stratified jitter,
forward + backward through `Network._batchify_rays` (value) and through `Network.forward` with host buffers
(e2e).  N > 1: every rank renders its own 6144 rays (weak scaling) and the gradients that leave the path
(hash table, MLP, point_dist, weight volume) are all-reduced with NCCL; no data-path collective.

`--impl reference` times the reference's algorithm on the host cores: the oracle restatement
(oracle/occnerf_oracle.py, CPU torch fp32 + the C hash grid; the reference itself cannot be imported on the
GPU box), forward + backward on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S_SAMPLES = 128
RAYS_PER_STEP = 6 * 32 * 32
FLOP_MLP_FWD = 923136.0          # per sample
FLOP_MLP_FWD_BWD = 2769408.0


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm / CPU baseline
def cpu_reference(sample_rays: int, steps: int, warmup: int, seed: int = 0, workload: str = "zju387"):
    """Forward + backward of the oracle restatement on `sample_rays` rays of the bench workload, all host cores."""
    from occnerf_b200 import synthetic as S
    from oracle import hashgrid_c, make_golden, occnerf_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    hashgrid_c.set_threads(cores)
    spec = WORKLOADS[workload]
    sub = S.make_subject(seed=0, bbox_offset=spec["bbox_offset"])
    w = S.make_weights(sub.bound, seed=0)
    fr = S.make_frame(sub, mode="patch", n_patches=6, patch=32, seed=seed, bbox_offset=spec["bbox_offset"], occlusion_band=spec["occlusion_band"])
    vol = S.make_motion_weights_vol(sub.priors, seed=0).requires_grad_(True)
    for t in [w.embeddings, sub.point_dist, w.geo_w, w.geo_b, w.out_w, w.out_b] + w.pts_w + w.pts_b + w.rgb_w + w.rgb_b:
        t.requires_grad_(True)
    n = min(sample_rays, fr.rays_o.shape[0])
    import dataclasses
    sl = {f.name: (getattr(fr, f.name)[:n] if f.name in ("rays_o", "rays_d", "near", "far") else getattr(fr, f.name))
          for f in dataclasses.fields(fr)}
    frs = S.Frame(**sl)
    t_rand = torch.rand(n, S_SAMPLES, generator=torch.Generator().manual_seed(1))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        out = O.render_rays(frs, vol, sub, w, iter_val=500, training=True, t_rand=t_rand)
        make_golden.scalar_loss(out).backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        for t in [vol, w.embeddings]:
            t.grad = None
    hashgrid_c.set_threads(1)
    ms = 1e3 * float(np.mean(times))
    return {"value": n / (ms / 1e3), "ms_per_step": ms, "cores": cores, "sample": f"{n} of {RAYS_PER_STEP} rays x {S_SAMPLES} samples, fwd+bwd, {steps} timed repeats"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    r = cpu_reference(args.ref_rays, max(1, args.steps), max(0, min(args.warmup, 1)), workload=args.workload)
    line = {"impl": "reference", "metric": "rays_per_sec_fwd_bwd_128spr", "value": r["value"], "unit": "rays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload]["name"], "rays_per_step": args.ref_rays, "samples_per_ray": S_SAMPLES},
            "cpu_baseline": {"value": r["value"], "unit": "rays/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- our arm
WORKLOADS = {
    # BASELINE configs[1]: ZJU-Mocap-387-shaped training step
    "zju387": dict(name="zju387_train_step_6x32x32_rays_128spr_fwd_bwd", bbox_offset=0.3, occlusion_band=None),
    # BASELINE configs[3]: OcMotion-shaped occluded training (configs/occnerf/ocmotion/*/occnerf.yaml: bbox_offset 2.0; the dataset
    # blanks a column band of the alpha mask, train.py:286-287; completeness loss + visibility counter update as in every training step)
    "ocmotion": dict(name="ocmotion_shaped_occluded_train_step_6x32x32_rays_128spr_fwd_bwd", bbox_offset=2.0, occlusion_band=(256, 60)),
}


class _WithActive:
    """SwitchReducer called with the structural-zero compaction map of a particular gradient list."""
    capturable = True

    def __init__(self, reducer, active):
        self.reducer, self.active = reducer, active

    def __call__(self, grads, hits=None):
        keep, self.reducer.active = self.reducer.active, self.active
        try:
            self.reducer(grads, hits=hits)
        finally:
            self.reducer.active = keep


class Workload:
    def __init__(self, device, rank, engine, workload="zju387", frame_rank=None):
        # frame_rank: which synthetic frame (patch placement) this rank renders.  Weak scaling keeps the per-GPU work FIXED: by
        # default every rank renders the frame of rank 0 -- the workload of the one-GPU line -- with its own stratified jitter and its
        # own targets, so the gradients differ and the all-reduce is real.  With rank-specific frames (--rank-frames distinct) the
        # step time of the eight frames spreads over 7.72 .. 8.22 ms on ONE GPU (profiles/r02_frame_variance.json) and a synchronous
        # data-parallel step follows the slowest frame: that is workload imbalance, not collective cost.
        frame_rank = rank if frame_rank is None else frame_rank
        from occnerf_b200 import synthetic as S
        from occnerf_b200.network import RenderConfig
        self.S = S
        self.spec = WORKLOADS[workload]
        sub = S.make_subject(seed=0, bbox_offset=self.spec["bbox_offset"])
        w = S.make_weights(sub.bound, seed=0)
        self.sub = sub
        self.net = S.network_from_synthetic(sub, w, RenderConfig(perturb=1.0, mlp_engine=engine), device=device)
        self.net.train(True)
        self.net.install_prologue()
        self.fr_host = S.make_frame(sub, mode="patch", n_patches=6, patch=32, seed=100 + frame_rank, bbox_offset=self.spec["bbox_offset"],
                                    occlusion_band=self.spec["occlusion_band"])
        self.fr = S.frame_to(self.fr_host, device)
        self.vol = S.make_motion_weights_vol(sub.priors, seed=0).to(device).requires_grad_(True)
        self.priors = sub.priors.to(device)
        gen = torch.Generator().manual_seed(7 + rank)
        self.target = torch.rand(RAYS_PER_STEP, 3, generator=gen).to(device)
        self.device = device
        self.iter_val = 500
        self.emb_fn, _ = self.net.get_non_rigid_embedder(6, 0, self.iter_val)
        self.packed = torch.cat([self.fr.rays_o, self.fr.rays_d, self.fr.near, self.fr.far], -1).contiguous()
        path_params = [p for n, p in self.net.named_parameters() if p.requires_grad and not n.startswith(("mweight_vol_decoder", "pose_decoder", "non_rigid_mlp"))]
        self.path_params = path_params
        from occnerf_b200.optim import ClipAdam
        from occnerf_b200.distributed import GradReducer
        self.opt_path = ClipAdam(path_params, lr=5e-4, max_norm=1.0)          # native clip + Adam (csrc/optim.cu), trainer.py:248-249
        self.opt_all = ClipAdam([p for p in self.net.parameters() if p.requires_grad], lr=5e-4, max_norm=1.0)
        from occnerf_b200.distributed import structural_zero_slices, SwitchReducer
        all_params = [p for p in self.net.parameters() if p.requires_grad]
        active_e2e = structural_zero_slices(all_params)
        self.emb = self.net.cnl_mlp.module.encoder.embeddings
        self.reducer_kind = "none"
        import torch.distributed as dist
        if dist.is_initialized() and dist.get_world_size() > 1:
            self.reducer_kind = "nccl"
            self.reducer = GradReducer()
            self.reducer_e2e = GradReducer(active=active_e2e)
            if os.environ.get("OCCNERF_REDUCER", "switch") == "switch":
                # gradient all-reduce as one of our kernels over NVSwitch peer memory (csrc/collective.cu); NCCL only if the
                # symmetric-memory rendezvous is not available on this box (every rank takes the same branch)
                ok = 1
                try:
                    # the decoder's transposed-convolution weight gradients (layers 2..5; the first keeps its compacted exchange) are
                    # written by occnerf_deconv3d_backward straight into the all-reduce buffer
                    dec = self.net.mweight_vol_decoder
                    convs = [m for m in dec.decoder.block_conv if isinstance(m, torch.nn.ConvTranspose3d)]
                    inplace = convs[1:] if dec.native else []
                    skip = [self.emb] + [c.weight for c in inplace]
                    bucket = sum((p[active_e2e[i]].numel() if i in active_e2e else p.numel()) for i, p in enumerate(all_params)
                                 if not any(p is q for q in skip))
                    bucket += self.vol.numel() + 6890 + 64
                    sw = SwitchReducer(self.emb.numel(), bucket, device, active=None, reserve_numel=sum(c.weight.numel() + 4 for c in inplace))
                    if inplace:
                        from occnerf_b200.prologue import _DecoderFn
                        dsts = [None] * 12
                        for c in inplace:
                            dsts[2 + 2 * convs.index(c)] = sw.reserve(tuple(c.weight.shape))
                        _DecoderFn.grad_out = dsts
                except Exception as exc:
                    print(f"bench.py: SwitchReducer unavailable ({type(exc).__name__}: {exc}); using NCCL", file=sys.stderr)
                    ok = 0
                flag = torch.tensor([ok], device=device)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                if int(flag.item()) == 1:
                    self.net.bind_emb_grad(sw.table_view)
                    self.reducer = sw
                    self.reducer_e2e = _WithActive(sw, active_e2e)
                    self.reducer_kind = sw.kind
        else:
            self.net.bind_emb_grad()                         # one persistent table-gradient buffer: a memset and an add less per step
        # pinned host copies of everything `Network.forward` receives per frame (trainer.py:223-229)
        h = self.fr_host
        self.host = {k: v.pin_memory() for k, v in dict(rays_o=h.rays_o, rays_d=h.rays_d, near=h.near, far=h.far, dst_Rs=h.dst_Rs,
                     dst_Ts=h.dst_Ts, cnl_gtfms=h.cnl_gtfms, priors=sub.priors, posevec=h.dst_posevec, bmin=h.cnl_bbox_min_xyz,
                     bscale=h.cnl_bbox_scale_xyz, bg=h.bgcolor, target=self.target.cpu()).items()}
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in self.host.values())
        self.loss_host = torch.zeros(1).pin_memory()

    def loss(self, out, target, world=1):
        """trainer.py:135-189 without the LPIPS term: 0.2 * img2mse(_unpack_imgs(rgb, ...), targets) + mean(comp_loss), value and
        gradients by occnerf_patch_loss (csrc/loss.cu).  The six 32 x 32 patches of the synthetic frame are fully covered by rays, so the
        unpacked images are the rays in patch order.  (data parallel: 1 / world folded into the weights, so that the SUM all-reduce of
        the gradients is their average)"""
        from occnerf_b200 import ops
        if not hasattr(self, "_patch_meta"):
            dev = out["rgb"].device
            n_patch = RAYS_PER_STEP // 1024
            self._patch_meta = (torch.ones(n_patch, 32, 32, device=dev, dtype=torch.uint8),
                                torch.arange(0, RAYS_PER_STEP + 1, 1024, device=dev, dtype=torch.int32), torch.zeros(3, device=dev))
        masks, div, bg = self._patch_meta
        return ops.patch_loss(out["rgb"], out["comp_loss"], masks, div, bg, target.reshape(-1, 32, 32, 3), 0.2 / world, 1.0 / world)[0]

    def grads_to_reduce(self):
        return [p.grad for p in self.path_params if p.grad is not None] + ([self.vol.grad] if self.vol.grad is not None else [])

    def step_device(self, world):
        """One step with inputs resident in HBM: the ray path only (vol is a leaf)."""
        net, fr = self.net, self.fr
        net.zero_bound_grads()
        out = net._batchify_rays(self.packed, pos_embed_fn=None, non_rigid_pos_embed_fn=self.emb_fn, non_rigid_mlp_input=None,
                                 motion_scale_Rs=fr.motion_scale_Rs[None], motion_Ts=fr.motion_Ts[None], motion_weights_vol=self.vol,
                                 cnl_bbox_min_xyz=fr.cnl_bbox_min_xyz, cnl_bbox_scale_xyz=fr.cnl_bbox_scale_xyz, bgcolor=fr.bgcolor)
        self.loss(out, self.target, world).backward()
        net.attach_bound_grads()
        hits = out["hits"]
        if world > 1:
            # one flat bucket for the 22 small tensors (MLP, point_dist, weight volume), the 59 MiB table gradient in place
            self.reducer([p.grad for p in self.path_params] + [self.vol.grad], hits=hits)
        self.opt_path.step()                             # occnerf_clip_adam_step: global-norm clip + Adam, 3 launches
        self.opt_path.zero_grad(set_to_none=True)
        self.vol.grad = None
        net.apply_visibility(hits)

    def make_graphed(self, world):
        """The e2e step as one CUDA graph (occnerf_b200/train_step.py); under data parallelism the NCCL all-reduces of the
        gradients and of the visibility votes are part of the graph."""
        from occnerf_b200 import distributed as D
        from occnerf_b200.train_step import GraphedTrainStep
        params = [p for p in self.net.parameters() if p.requires_grad]
        sync = self.reducer_e2e if world > 1 else None       # (SwitchReducer is capturable: one graph; NCCL: two graphs around it)
        self.graphed = GraphedTrainStep(self.net, self.opt_all, lambda out, d: self.loss({"rgb": out["rgb"], "comp_loss": out["comp_loss"]}, d["target"], world),
                                        self.host, self.iter_val, params=params, max_norm=None, grad_sync=sync)
        return self.graphed

    def step_e2e_graphed(self, world):
        self.graphed.step(self.host)

    def step_e2e(self, world):
        """One step through the public API with host buffers: H2D of the frame, Network.forward, loss, backward,
        all-reduce, optimizer, D2H of the loss."""
        d = {k: v.to(self.device, non_blocking=True) for k, v in self.host.items()}
        net = self.net
        net.zero_bound_grads()
        out = net.forward((d["rays_o"], d["rays_d"]), d["dst_Rs"], d["dst_Ts"], d["cnl_gtfms"], d["priors"], dst_posevec=d["posevec"],
                          near=d["near"], far=d["far"], iter_val=self.iter_val, cnl_bbox_min_xyz=d["bmin"], cnl_bbox_scale_xyz=d["bscale"],
                          bgcolor=d["bg"])
        loss = self.loss({"rgb": out["rgb"], "comp_loss": out["comp_loss"]}, d["target"], world)
        loss.backward()
        net.attach_bound_grads()
        params = [p for p in net.parameters() if p.requires_grad]
        if world > 1:
            self.reducer_e2e([p.grad for p in params], hits=out["hits"])
        self.opt_all.step()
        self.opt_all.zero_grad(set_to_none=True)
        net.apply_visibility(out["hits"])
        self.loss_host.copy_(loss.detach().reshape(1), non_blocking=True)


def forward_frame(wl, flush, repeats=3):
    """BASELINE configs[0] shape on the GPU: forward render of every bbox-hitting pixel of one 512x512 view (eval mode, no
    jitter, no gradients), through `Network._batchify_rays` in chunks of 32 768 rays.  Reported beside the training metric."""
    S = wl.S
    fr = S.frame_to(S.make_frame(wl.sub, mode="full", img=512, seed=3), wl.device)
    packed = torch.cat([fr.rays_o, fr.rays_d, fr.near, fr.far], -1).contiguous()
    net = wl.net
    was_training, perturb = net.training, net.cfg.perturb
    net.train(False)
    net.cfg.perturb = 0.0
    emb_fn, _ = net.get_non_rigid_embedder(6, 0, 10 ** 7)
    vol = wl.vol.detach()

    def render():
        with torch.no_grad():
            return net._batchify_rays(packed, pos_embed_fn=None, non_rigid_pos_embed_fn=emb_fn, non_rigid_mlp_input=fr.dst_posevec[None],
                                      motion_scale_Rs=fr.motion_scale_Rs[None], motion_Ts=fr.motion_Ts[None], motion_weights_vol=vol,
                                      cnl_bbox_min_xyz=fr.cnl_bbox_min_xyz, cnl_bbox_scale_xyz=fr.cnl_bbox_scale_xyz, bgcolor=fr.bgcolor)
    render()
    torch.cuda.synchronize()
    from occnerf_b200 import _lib
    _lib.PROFILE = {}
    render()
    torch.cuda.synchronize()
    prof, _lib.PROFILE = _lib.PROFILE, None
    top = sorted(((sum(a.elapsed_time(b) for a, b, _w in evs), name) for name, evs in prof.items()), reverse=True)[:6]
    tot = 0.0
    for _ in range(repeats):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = render()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    net.train(was_training)
    net.cfg.perturb = perturb
    ms = tot / repeats
    n = packed.shape[0]
    peaks = load_peaks()
    t_mlp = dict((name, t) for t, name in top).get("occnerf_mlp_forward_tc")
    roof = None
    if t_mlp:
        ach = FLOP_MLP_FWD * n * S_SAMPLES / (t_mlp * 1e-3) / 1e12
        roof = {"kernel": "occnerf_mlp_forward_tc", "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_tflops_sustained"], "traffic": None}
    return {"workload": "forward render of one 512x512 view, all bbox-hitting pixels, 128 samples/ray, non-rigid MLP active (iter 1e7)",
            "rays": n, "ms_per_frame": ms, "rays_per_sec": n / (ms * 1e-3), "finite": bool(torch.isfinite(out["rgb"]).all()),
            "top_calls_ms": [[name, round(t, 2)] for t, name in top], "roofline": roof}


def cpu_forward_reference(sample_rays: int):
    """BASELINE configs[0]'s own measurement: the reference's pure-PyTorch forward on the host cores (oracle port), bounded sample."""
    from occnerf_b200 import synthetic as S
    from oracle import hashgrid_c, occnerf_oracle as O
    import dataclasses
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    hashgrid_c.set_threads(cores)
    sub = S.make_subject(seed=0)
    w = S.make_weights(sub.bound, seed=0)
    fr = S.make_frame(sub, mode="full", img=512, seed=3)
    vol = S.make_motion_weights_vol(sub.priors, seed=0)
    pick = torch.arange(0, fr.rays_o.shape[0], max(1, fr.rays_o.shape[0] // sample_rays))[:sample_rays]
    sl = {f.name: (getattr(fr, f.name)[pick] if f.name in ("rays_o", "rays_d", "near", "far") else getattr(fr, f.name)) for f in dataclasses.fields(fr)}
    frs = S.Frame(**sl)
    times = []
    with torch.no_grad():
        for it in range(3):
            t0 = time.perf_counter()
            O.render_rays(frs, vol, sub, w, iter_val=10 ** 7, training=False)
            times.append(time.perf_counter() - t0)
    hashgrid_c.set_threads(1)
    t = float(np.mean(times[1:]))
    n = int(pick.numel())
    return {"value": n / t, "unit": "rays/s", "cores": cores, "kind": "port", "sample": f"{n} of {fr.rays_o.shape[0]} rays of the frame x {S_SAMPLES} samples, forward only, 2 timed repeats"}


def timed_loop(fn, steps, warmup, world, flush):
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        flush.fill_(1.0)                         # evict L2 (126 MB) between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps


def kernel_table(profile, steps):
    rows = []
    for name, evs in profile.items():
        ms = sum(a.elapsed_time(b) for a, b, _w in evs)
        rows.append({"call": name, "ms_per_step": ms / steps, "launches_per_step": len(evs) / steps, "work_per_step": sum(w for _a, _b, w in evs) / steps})
    rows.sort(key=lambda r: -r["ms_per_step"])
    return rows


# DRAM bytes per sample (dram__bytes_read.sum + dram__bytes_write.sum of one launch / samples of that launch) from the
# `ncu --set full` capture summarised in profiles/r01c_kernels_full.md (300 000-sample launches of the same command)
NCU_TRAFFIC_PER_SAMPLE = {"occnerf_mlp_forward_tc": (0.151179e9 + 1.470747e9) / 300000, "occnerf_mlp_backward_tc": (0.128911e9 + 1.335716e9) / 300000,
                          "occnerf_mlp_wgrad_tc": (2.805448e9 + 0.006918e9) / 300000, "occnerf_aggregate_backward": (0.194468e9 + 0.015459e9) / 300000,
                          "occnerf_aggregate_forward": (0.054993e9 + 0.052413e9) / 300000, "occnerf_hashgrid_backward": (0.113153e9 + 0.007391e9) / 300000}
MMA_ISSUE_FACTOR = {"tc3": 3.0, "tc3b1": 3.0, "tc1": 1.0, "tf32": 2.0}   # bf16-equivalent tensor-pipe work per algorithmic product (split-bf16: 3 MMAs; tf32: half rate)


def roofline_for(row, peaks, M, engine="tc3"):
    """Algorithmic work of one C call (SURVEY.md 8d / DESIGN.md) divided by its measured device time."""
    name, ms = row["call"], row["ms_per_step"] / max(row["launches_per_step"], 1e-9)
    per_launch_samples = M / max(row["launches_per_step"], 1.0) if name.startswith(("occnerf_mlp", "occnerf_aggregate", "occnerf_hashgrid", "occnerf_sample")) else M
    traffic = NCU_TRAFFIC_PER_SAMPLE.get(name)
    traffic = traffic * per_launch_samples if traffic is not None else None
    hbm = {"occnerf_warp_forward": 20.25, "occnerf_warp_backward": 4.0, "occnerf_warp_forward_packed": 20.25, "occnerf_warp_backward_packed": 4.0,
           # weight gradients: the bf16 operands read once = (8 x 256 + 80 + 144) + (8 x 256 + 80 + 16) columns x 2 B per sample
           "occnerf_mlp_wgrad_tc": 8832.0, "occnerf_composite_forward": 28.0 + 28.0 / S_SAMPLES,
           "occnerf_composite_backward": 52.0, "occnerf_hashgrid_forward": 144.0, "occnerf_hashgrid_backward": 128.0,
           "occnerf_aggregate_forward": 160.0 + 144.0, "occnerf_aggregate_backward": 160.0 + 144.0, "occnerf_knn": 12.0 + 160.0,
           "occnerf_knn_grid": 12.0 + 160.0, "occnerf_knn_tree": 12.0 + 160.0, "occnerf_sample_geometry": 12.0 + 40.0 + 20.0}
    if name in ("occnerf_sgemm", "occnerf_mlp_forward_tc", "occnerf_mlp_backward_tc"):
        flops = row["work_per_step"] / max(row["launches_per_step"], 1e-9)
        ach = flops / (ms * 1e-3) / 1e12
        out = {"kernel": name, "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
               "frac": ach / peaks["bf16_tflops_sustained"], "traffic": traffic, "peak_source": peaks["source"] + " bf16 dense (sustained)"}
        if name in ("occnerf_mlp_forward_tc", "occnerf_mlp_backward_tc") and engine in MMA_ISSUE_FACTOR:
            # what the tensor pipe actually executes: every fp32-grade product is 3 bf16 MMAs in the split-bf16 engine
            out["issued_mma_tflops"] = ach * MMA_ISSUE_FACTOR[engine]
            out["issued_mma_frac"] = out["issued_mma_tflops"] / peaks["bf16_tflops_sustained"]
        return out
    if name == "occnerf_clip_adam_step":                  # 32 B per parameter (g twice, p/m/v read + written)
        bytes_ = 32.0 * row.get("params", 0.0)
        ach = bytes_ / (ms * 1e-3) / 1e9
        return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                "traffic": None, "peak_source": peaks["source"] + " copy bandwidth"}
    bytes_ = hbm.get(name, 0.0) * per_launch_samples
    ach = bytes_ / (ms * 1e-3) / 1e9
    out = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
           "traffic": traffic, "peak_source": peaks["source"] + " copy bandwidth"}
    # The gather / scatter stages keep their tables in L2 by design (BASELINE north star: "HBM/L2 GB/s for the gather stages"): their
    # algorithmic GATHERED bytes per sample (the packed warp kernel reads 24 32-byte cells) against the L2 output cap of
    # /opt/skills/guides/B300_MICROARCH.md (~6300 B/clk over the chip at the SM clock), next to the HBM figure above.
    l2 = {"occnerf_warp_forward_packed": 768.0, "occnerf_hashgrid_forward": 2048.0, "occnerf_hashgrid_backward": 2048.0,
          "occnerf_aggregate_forward": 40 * 144.0, "occnerf_aggregate_backward": 40 * 144.0}
    if name in l2:
        cap = 6300.0 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e9
        g = l2[name] * per_launch_samples / (ms * 1e-3) / 1e9
        out["l2"] = {"gathered_bytes_per_sample": l2[name], "achieved_gbs": g, "cap_gbs": cap, "frac": g / cap,
                     "cap_source": "B300_MICROARCH.md LTS throughput cap 6300 B/clk x SM clock"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default=os.environ.get("OCCNERF_ENGINE", "tf32"), choices=["fp32", "tf32", "tc3", "tc3b1", "tc1"])
    ap.add_argument("--workload", default="zju387", choices=list(WORKLOADS), help="zju387 = BASELINE configs[1] (the headline); ocmotion = configs[3]")
    ap.add_argument("--ref-rays", type=int, default=1024, help="rays per step of the CPU reference arm / cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time the steps eagerly instead of as CUDA graphs")
    ap.add_argument("--rank-frames", default="same", choices=["same", "distinct"],
                    help="N > 1: every rank renders the configs[1] frame of the one-GPU line (fixed per-GPU work; default) or its own frame")
    ap.add_argument("--profile-mode", action="store_true", help="device-resident steps only (no e2e, no CPU baseline): for ncu runs")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback; use --impl reference for the CPU arm)")
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout for the one JSON line
        dist.init_process_group("nccl", device_id=device)
    from occnerf_b200 import _lib
    _lib.load()
    torch.manual_seed(1234 + rank)                   # rank-specific stratified jitter (torch.rand inside the path)
    wl = Workload(device, rank, args.engine, args.workload, frame_rank=0 if args.rank_frames == "same" else rank)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=device)
    M = RAYS_PER_STEP * S_SAMPLES

    # ---- value: ray path, inputs resident.  The step is captured ONCE as a CUDA graph and replayed (the data-parallel all-reduce is
    # one of our kernels and is captured with it); an eager pass with per-call CUDA events feeds the kernel table and the launch count.
    sampler = ClockSampler(local)
    sampler.start()
    _lib.PROFILE = {}
    c0 = dict(_lib.COUNTERS)
    # the per-call table (and `ms_per_step_eager`) is measured with the weight-gradient kernel in line on the one stream, so that the
    # figures of the calls it otherwise overlaps (hash-grid / aggregation backward) stay per-kernel times; the timed graph has the overlap
    from occnerf_b200 import mlp_tc as _mlp_tc
    overlap_default, _mlp_tc.WGRAD_OVERLAP = _mlp_tc.WGRAD_OVERLAP, False
    for _ in range(args.warmup):
        wl.step_device(world)
    _lib.PROFILE, c0 = {}, dict(_lib.COUNTERS)
    if args.profile_mode:
        torch.cuda.profiler.start()              # `ncu --profile-from-start off` then sees the steady-state steps only
    ms_eager = timed_loop(lambda: wl.step_device(world), args.steps, 0, world, flush)
    if args.profile_mode:
        torch.cuda.profiler.stop()
    _mlp_tc.WGRAD_OVERLAP = overlap_default
    value_launch = "eager"
    ms = ms_eager
    if not args.no_graph and not args.profile_mode and getattr(getattr(wl, "reducer", None), "capturable", world == 1):
        ok = 1
        try:
            prof_keep, _lib.PROFILE = _lib.PROFILE, None
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                wl.step_device(world)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g_value = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_value):
                wl.step_device(world)
            _lib.PROFILE = prof_keep
        except Exception as exc:
            import traceback
            traceback.print_exc(file=sys.stderr)
            print(f"bench.py: CUDA-graph capture of the device-resident step failed ({type(exc).__name__}: {exc}); keeping the eager timing", file=sys.stderr)
            ok = 0
        if world > 1:
            import torch.distributed as dist
            flag = torch.tensor([ok], device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = int(flag.item())
        if ok:
            prof_keep, _lib.PROFILE = _lib.PROFILE, None
            ms = timed_loop(lambda: g_value.replay(), args.steps, args.warmup, world, flush)
            _lib.PROFILE = prof_keep
            value_launch = "cuda_graph"
    if world > 1 and wl.reducer_kind != "nccl":
        import ctypes
        buf = (ctypes.c_ulonglong * 4)()
        _lib.load().occnerf_allreduce_debug(ctypes.cast(buf, ctypes.c_void_p), 1)
        n_ar = max(int(buf[3]), 1)
        ar_phases = {"wait_arrive_us": buf[0] / n_ar / 1e3, "data_us": buf[1] / n_ar / 1e3, "wait_finish_us": buf[2] / n_ar / 1e3, "launches": int(buf[3])}
        print(f"bench.py: rank {rank} all-reduce phases per launch {ar_phases}", file=sys.stderr)
    launches = (_lib.COUNTERS["launches"] - c0["launches"]) / args.steps
    profile, _lib.PROFILE = _lib.PROFILE, None
    clocks = sampler.stop()
    table = kernel_table(profile, args.steps)
    n_path_params = float(sum(p.numel() for p in wl.path_params))
    for r in table:
        if r["call"] == "occnerf_clip_adam_step":
            r["params"] = n_path_params
    # ---- e2e: public API with host buffers
    if args.profile_mode:
        if rank == 0:
            print(json.dumps({"profile_mode": True, "ms_per_step": ms, "kernels": [(r["call"], round(r["ms_per_step"], 4)) for r in table]}))
        return
    e2e_mode, e2e_launches = "eager", None
    # one CUDA graph per step on one GPU; with data parallelism two graphs (forward+backward | clip+Adam) around the eagerly
    # launched NCCL all-reduces -- the same launch mode at every rank count
    if not args.no_graph:
        ok = 1
        try:
            g = wl.make_graphed(world)
            e2e_mode, e2e_launches = ("cuda_graph" if g.graph_opt is None else "2 cuda graphs + eager nccl"), g.launches
        except Exception as exc:                                   # capture is an optimisation, never a requirement
            import traceback
            traceback.print_exc(file=sys.stderr)
            print(f"bench.py: CUDA-graph capture of the e2e step failed ({type(exc).__name__}: {exc}); timing the eager step", file=sys.stderr)
            ok = 0
        if world > 1:                                              # every rank replays, or none does
            import torch.distributed as dist
            flag = torch.tensor([ok], device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                e2e_mode, e2e_launches = "eager", None
    step = wl.step_e2e_graphed if e2e_mode != "eager" else wl.step_e2e
    ms_e2e = timed_loop(lambda: step(world), args.steps, args.warmup, world, flush)
    # the same K steps by the wall clock (host gaps between steps included; L2 flushes included too, hence a little above the event time)
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        step(world)
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t_wall) * 1e3 / args.steps

    if rank == 0:
        peaks = load_peaks()
        roof_all = [roofline_for(r, peaks, M, args.engine) for r in table[:10]]
        line = {
            "metric": "rays_per_sec_fwd_bwd_128spr", "value": world * RAYS_PER_STEP / (ms * 1e-3), "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "tf32": "tf32->f32", "tc3": "bf16x3(split)->f32", "tc3b1": "bf16x3(split)->f32 fwd, bf16->f32 dgrad", "tc1": "bf16->f32"}[args.engine], "data": "synthetic",
            "config": {"workload": wl.spec["name"], "rays_per_step_per_gpu": RAYS_PER_STEP, "samples_per_ray": S_SAMPLES,
                       "mlp_engine": args.engine, "l2": "256 MiB flush write between timed steps", "timing": "cuda events per step, max over ranks",
                       "optimizer": "global-norm clip + Adam inside the step (occnerf_clip_adam_step)", "parallelism": f"dp{world}", "grad_allreduce": wl.reducer_kind,
                       "launch": value_launch, "ms_per_step_eager": ms_eager,
                       "rank_frames": args.rank_frames if world > 1 else "n/a",
                       "knn": "exact; pykeops' own reduction is un-vendored upstream (parity unpinned for that one call), ids checked against brute force"},
            "clocks": clocks,
            "e2e": {"value": world * RAYS_PER_STEP / (ms_e2e * 1e-3), "unit": "rays/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": wl.h2d_bytes,
                    "d2h_bytes_per_step": 4, "api": "Network.forward (prologue + ray path) + loss + backward + optimizer",
                    "launch": e2e_mode, "our_launches_per_step": e2e_launches, "wall_ms_per_step": wall_ms,
                    "wall_note": "K back-to-back steps by the host clock, no L2 flush between them"},
            "gpu_launches": launches,
            "roofline": roof_all[0] if roof_all else None,
            "kernels": [{"call": r["call"], "ms_per_step": round(r["ms_per_step"], 4), "launches_per_step": r["launches_per_step"]} for r in table],
            "roofline_all": roof_all,
        }
        if world == 1:
            line["forward_only"] = forward_frame(wl, flush)
            if not args.no_cpu_baseline:
                line["forward_only"]["cpu_baseline"] = cpu_forward_reference(args.ref_rays)
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference(args.ref_rays, 2, 1, workload=args.workload)
            line["cpu_baseline"] = {"value": cb["value"], "unit": "rays/s", "cores": cb["cores"], "kind": "port", "sample": cb["sample"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
