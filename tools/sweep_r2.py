"""Stress sweep of the WHOLE path (BASELINE.json configs[4]): 2^16 .. 2^22 rays x 64 / 128 / 256 samples per ray, 16-level hash
grid with a 2^19 table, forward-only and forward+backward, rays sharded over the ranks (strong scaling: N rays in total).

    python tools/sweep_r2.py [--max-log2 22] [--engine tf32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/sweep_r2.py

Rays: the 6144 patch rays of the bench frame tiled with a 1 % direction jitter (they all cross the body's box).  A batch is
processed in chunks of 32 768 rays (cfg.chunk) -- forward: `_batchify_rays` in eval mode; forward+backward: training mode,
`backward()` per chunk (gradients accumulate), one gradient all-reduce at the end when world > 1.  Timing: CUDA events around the
whole batch after one warm-up batch, max over ranks.  Per size the table also lists the time of the dominant C calls (MLP chains,
weight gradients, aggregation, hash grid, KNN) from the per-call events.  Output: markdown on stdout (committed under profiles/)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--min-log2", type=int, default=16)
ap.add_argument("--max-log2", type=int, default=22)
ap.add_argument("--engine", default="tf32")
ap.add_argument("--samples", default="64,128,256")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
d = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=d)
from occnerf_b200 import _lib, synthetic as S, distributed as D
from occnerf_b200.network import RenderConfig

sub = S.make_subject(0)
w = S.make_weights(sub.bound, seed=0)
vol = S.make_motion_weights_vol(sub.priors, 0).to(d)
fr0 = S.frame_to(S.make_frame(sub, mode="patch", n_patches=6, patch=32, seed=100), d)
CALLS = ["occnerf_mlp_forward_tc", "occnerf_mlp_backward_tc", "occnerf_mlp_wgrad_tc", "occnerf_aggregate_forward", "occnerf_aggregate_backward",
         "occnerf_hashgrid_forward", "occnerf_hashgrid_backward", "occnerf_knn_grid"]
rows = []
for Sn in [int(s) for s in args.samples.split(",")]:
    net = S.network_from_synthetic(sub, w, RenderConfig(perturb=1.0, mlp_engine=args.engine, N_samples=Sn), device=d)
    emb_fn, _ = net.get_non_rigid_embedder(6, 0, 500)
    params = [p for p in net.parameters() if p.requires_grad]
    reducer = D.GradReducer()
    for logn in range(args.min_log2, args.max_log2 + 1, 2):
        N = (2 ** logn) // world
        gen = torch.Generator(device=d).manual_seed(logn * 31 + rank)
        idx = torch.randint(0, fr0.rays_o.shape[0], (N,), device=d, generator=gen)
        jit = 1.0 + 0.01 * torch.randn(N, 3, device=d, generator=gen)
        rays = torch.cat([fr0.rays_o[idx], fr0.rays_d[idx] * jit, fr0.near[idx], fr0.far[idx]], -1).contiguous()
        kw = dict(pos_embed_fn=None, non_rigid_pos_embed_fn=emb_fn, non_rigid_mlp_input=None, motion_scale_Rs=fr0.motion_scale_Rs[None],
                  motion_Ts=fr0.motion_Ts[None], cnl_bbox_min_xyz=fr0.cnl_bbox_min_xyz, cnl_bbox_scale_xyz=fr0.cnl_bbox_scale_xyz, bgcolor=fr0.bgcolor)

        def forward_only():
            net.train(False)
            net.cfg.perturb = 0.0
            with torch.no_grad():
                for i in range(0, N, 32768):
                    net._batchify_rays(rays[i:i + 32768], motion_weights_vol=vol, **kw)

        def forward_backward():
            net.train(True)
            net.cfg.perturb = 1.0
            v = vol.detach().requires_grad_(True)
            for i in range(0, N, 32768):
                out = net._batchify_rays(rays[i:i + 32768], motion_weights_vol=v, **kw)
                (out["rgb"].square().mean() + out["comp_loss"].mean()).backward()
            if world > 1:
                reducer([p.grad for p in params] + [v.grad])
            for p in params:
                p.grad = None

        for mode, fn in (("fwd", forward_only), ("fwd+bwd", forward_backward)):
            fn()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            _lib.PROFILE = {}
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            prof, _lib.PROFILE = _lib.PROFILE, None
            t = torch.tensor([e0.elapsed_time(e1)], device=d)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            calls = {c: sum(a.elapsed_time(b) for a, b, _w in prof.get(c, [])) for c in CALLS}
            rows.append(dict(log2_rays=logn, samples=Sn, mode=mode, n_gpus=world, ms=ms, rays_per_s=(2 ** logn) / (ms * 1e-3),
                             msamples_per_s=(2 ** logn) * Sn / (ms * 1e-3) / 1e6, calls_ms={k[8:]: round(v, 2) for k, v in calls.items() if v > 0}))
        del rays, idx, jit
        torch.cuda.empty_cache()
if rank == 0:
    print(f"Whole-path stress sweep, {world} x B200, engine {args.engine}, rays sharded over the ranks; per-call times are rank 0's.\n")
    print("| rays | samples/ray | mode | GPUs | ms | rays/s | Msamples/s | dominant calls (ms, rank 0) |")
    print("|---:|---:|---|---:|---:|---:|---:|---|")
    for r in rows:
        print(f"| 2^{r['log2_rays']} | {r['samples']} | {r['mode']} | {r['n_gpus']} | {r['ms']:.1f} | {r['rays_per_s']:.3e} | {r['msamples_per_s']:.1f} | "
              + ", ".join(f"{k} {v}" for k, v in r["calls_ms"].items()) + " |")
    print("\n" + json.dumps(rows))
if world > 1:
    dist.destroy_process_group()
