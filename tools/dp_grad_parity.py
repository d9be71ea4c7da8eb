"""Data-parallel gradient parity on real GPUs: N ranks each render their contiguous share of ONE ray batch,
the gradients are all-reduced by SwitchReducer (csrc/collective.cu; OCCNERF_REDUCER=nccl: GradReducer), and the result is compared
on every rank with a single-GPU backward over the whole batch.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dp_grad_parity.py
Prints one JSON object (rank 0): relative Frobenius errors per gradient tensor, exactness of the visibility votes."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from occnerf_b200 import synthetic as S
from occnerf_b200.network import RenderConfig
from occnerf_b200.distributed import SwitchReducer, GradReducer, shard_range

ENGINE = os.environ.get("ENGINE", "tf32")
PATCH = 32
sub = S.make_subject(seed=0)
w = S.make_weights(sub.bound, seed=0, table_scale=0.05, nonzero_bias=True)
w.geo_b[0] = 20.0                                      # dense densities: opaque rays, many visibility votes
fr = S.make_frame(sub, mode="patch", n_patches=2 * world, patch=PATCH, seed=4)
vol0 = S.make_motion_weights_vol(sub.priors, seed=0)
N = fr.rays_o.shape[0]
t_rand = torch.rand(N, 128, generator=torch.Generator().manual_seed(23))
target = torch.rand(N, 3, generator=torch.Generator().manual_seed(5))


def run(lo, hi, scale, reducer=None):
    net = S.network_from_synthetic(sub, w, RenderConfig(perturb=1.0, mlp_engine=ENGINE, knn_mode="grid"), device=dev)
    net.train(True)
    frd = S.frame_to(fr, dev)
    vol = vol0.to(dev).requires_grad_(True)
    emb_fn, _ = net.get_non_rigid_embedder(6, 0, 500)
    packed = torch.cat([frd.rays_o, frd.rays_d, frd.near, frd.far], -1)[lo:hi].contiguous()
    if reducer is not None and hasattr(reducer, "bind_table"):
        reducer.bind_table(net)
        reducer.zero_table()
    out = net._batchify_rays(packed, pos_embed_fn=None, non_rigid_pos_embed_fn=emb_fn, non_rigid_mlp_input=None,
                             motion_scale_Rs=frd.motion_scale_Rs[None], motion_Ts=frd.motion_Ts[None], motion_weights_vol=vol,
                             cnl_bbox_min_xyz=frd.cnl_bbox_min_xyz, cnl_bbox_scale_xyz=frd.cnl_bbox_scale_xyz, bgcolor=frd.bgcolor,
                             t_rand=t_rand[lo:hi].to(dev))
    loss = (0.2 * torch.mean((out["rgb"] - target[lo:hi].to(dev)) ** 2) + out["comp_loss"].mean()) * scale
    loss.backward()
    net.attach_bound_grads()
    params = [(n, p) for n, p in net.named_parameters() if p.grad is not None]
    grads = [p.grad for _, p in params] + [vol.grad]
    hits = out["hits"].clone()
    if reducer is not None:
        reducer(grads, hits=hits)
    torch.cuda.synchronize()
    return {**{n: p.grad.detach().clone() for n, p in params}, "motion_weights_vol": vol.grad.detach().clone()}, hits


# the whole batch on this GPU
full, hits_full = run(0, N, 1.0)
lo, hi = shard_range(N, rank, world, granule=PATCH * PATCH)
if os.environ.get("OCCNERF_REDUCER", "switch") == "switch":
    n_small = sum(g.numel() for k, g in full.items() if "embeddings" not in k) + 6890 + 64
    emb_numel = next(g.numel() for k, g in full.items() if "embeddings" in k)
    red = SwitchReducer(emb_numel, n_small, dev)
    kind = red.kind
else:
    red, kind = GradReducer(), "nccl"
part, hits_dp = run(lo, hi, 1.0 / world, red)
errs = {k: float((part[k].double() - full[k].double()).norm() / full[k].double().norm().clamp_min(1e-300)) for k in full}
res = {"world": world, "engine": ENGINE, "reducer": kind, "rays_total": N, "rays_per_rank": hi - lo,
       "rel_fro_err": errs, "worst": max(errs.values()),
       "votes_equal": bool(torch.equal(hits_dp.clamp(max=1.0), hits_full.clamp(max=1.0))), "votes": int(hits_full.sum().item())}
ok = torch.tensor([1.0 if (res["worst"] < float(os.environ.get("TOL", "2e-2")) and res["votes_equal"]) else 0.0], device=dev)
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
res["all_ranks_ok"] = bool(ok.item() == 1.0)
if rank == 0:
    print(json.dumps(res, indent=1))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if res["all_ranks_ok"] else 1)
