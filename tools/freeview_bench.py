"""BASELINE configs[2] as a runnable measurement (NOT part of bench.py's contract; written at the end of round 1, first
numbers are round-2 work): freeview render of V views at RES x RES, synthetic pose, the valid rays of every view sharded
over the ranks by contiguous range, no collective.  Per view and rank: rays generated on the device
(occnerf_generate_rays), rendered through Network.forward in eval mode, painted into an 8-bit frame on the device
(occnerf_unpack_image); the frame is read back to the host inside the timed region (that is the product of a render).

    python tools/freeview_bench.py --views 10 --res 1024                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
        tools/freeview_bench.py --views 10 --res 1024

Timing: CUDA events per view on the current stream, 2 warm-up views, max over ranks of the summed time; rank 0 prints
one JSON line (rays/s over all ranks = valid rays of all timed views / that time)."""
from __future__ import annotations

import argparse
import json
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=10)
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--engine", default="tf32")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("freeview_bench: needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from occnerf_b200 import render, synthetic as S
    from occnerf_b200.network import RenderConfig
    sub = S.make_subject(seed=0)
    w = S.make_weights(sub.bound, seed=0)
    net = S.network_from_synthetic(sub, w, RenderConfig(perturb=0.0, mlp_engine=args.engine), device=dev).train(False)
    net.install_prologue()
    fr_host = S.make_frame(sub, mode="patch", n_patches=1, patch=8, seed=9)       # only its pose / bbox fields are used
    fr = S.frame_to(fr_host, dev)
    data = dict(dst_Rs=fr.dst_Rs, dst_Ts=fr.dst_Ts, cnl_gtfms=fr.cnl_gtfms, motion_weights_priors=sub.priors.to(dev),
                dst_posevec=fr.dst_posevec, cnl_bbox_min_xyz=fr.cnl_bbox_min_xyz, cnl_bbox_scale_xyz=fr.cnl_bbox_scale_xyz,
                bgcolor=fr.bgcolor)
    box = {"min_xyz": np.array([-0.95, -1.35, -0.45], np.float32), "max_xyz": np.array([0.95, 0.65, 0.45], np.float32)}
    total = args.views + args.warmup
    frame_host = torch.empty(args.res, args.res, 3, dtype=torch.uint8).pin_memory()
    ms, rays = 0.0, 0
    for v in range(total):
        yaw = 2.0 * math.pi * v / max(1, args.views)                            # freeview.py:133-142: one turn about the vertical axis
        K, R, T = S.lookat_camera(args.res, yaw=yaw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        out = render.render_view(net, args.res, args.res, K, R.astype(np.float64), T.astype(np.float64), box, data, [0.0, 0.0, 0.0],
                                 rank=rank, world=world)
        frame_host.copy_(out["rgb8"], non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        if v >= args.warmup:
            ms += e0.elapsed_time(e1)
            rays += out["rays"][2]
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"metric": "rays_per_sec_fwd_128spr", "value": rays / (t.item() * 1e-3), "unit": "rays/s", "n_gpus": world,
                          "views": args.views, "warmup": args.warmup, "ms_per_view": t.item() / args.views, "scaling": "strong",
                          "config": {"workload": f"freeview_{args.views}x{args.res}x{args.res}_128spr_fwd", "valid_rays": rays,
                                     "mlp_engine": args.engine, "timing": "cuda events per view, max over ranks, frame read back inside"}}),
              flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
