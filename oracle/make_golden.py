"""ORACLE tooling: write tests/golden/*.npz by running the UNMODIFIED reference (through oracle/ref_shim.py).

Run in the build container only (needs /root/reference):
    python -m oracle.make_golden

Every case is regenerated from seeds by occnerf_b200/synthetic.py; the files hold the reference's outputs
(and the small inputs, so that a generator change is detected rather than silently shifting the target).
Large dense gradients (hash table, weight volume) are stored sparse.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occnerf_b200 import synthetic as S  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (subject/weights kwargs, frame kwargs, render kwargs)
    "train_dense": dict(weights=dict(seed=0, table_scale=0.05, nonzero_bias=True, sigma_bias=20.0),
                        frame=dict(mode="patch", n_patches=2, patch=8, seed=3),
                        render=dict(iter_val=150000, training=True, perturb=1.0), t_rand_seed=11),
    "eval_init": dict(weights=dict(seed=0, table_scale=1e-4, nonzero_bias=False, sigma_bias=0.0),
                      frame=dict(mode="image", img=128, max_rays=192, seed=5),
                      render=dict(iter_val=10000000, training=False, perturb=0.0), t_rand_seed=None),
    "train_init_early": dict(weights=dict(seed=1, table_scale=1e-4, nonzero_bias=False, sigma_bias=0.0),
                             frame=dict(mode="patch", n_patches=1, patch=8, seed=9),
                             render=dict(iter_val=500, training=True, perturb=1.0), t_rand_seed=13),
}

LOSS_W = dict(rgb=(1.0, 2.0, 3.0), alpha=0.5, depth=0.25, comp=1.0)


def scalar_loss(out):
    """A fixed linear functional of the outputs so that every output gets a non-trivial upstream gradient."""
    rgbw = torch.tensor(LOSS_W["rgb"], dtype=out["rgb"].dtype, device=out["rgb"].device)
    loss = (out["rgb"] * rgbw).sum() + out["alpha"].sum() * LOSS_W["alpha"] + out["depth"].sum() * LOSS_W["depth"]
    if out["comp_loss"].numel() > 1:
        loss = loss + out["comp_loss"].mean() * LOSS_W["comp"]
    return loss


def build_case(name):
    c = CASES[name]
    sub = S.make_subject(seed=0)
    wk = dict(c["weights"])
    sigma_bias = wk.pop("sigma_bias")
    w = S.make_weights(sub.bound, **wk)
    w.geo_b[0] = sigma_bias
    fr = S.make_frame(sub, **c["frame"])
    vol = S.make_motion_weights_vol(sub.priors, seed=0)
    t_rand = None
    if c["t_rand_seed"] is not None:
        t_rand = torch.rand(fr.rays_o.shape[0], 128, generator=torch.Generator().manual_seed(c["t_rand_seed"]))
    return sub, w, fr, vol, t_rand, c["render"]


def _sparse(t, stride=1):
    """(indices, values) of every `stride`-th non-zero entry -- keeps the fixtures small."""
    flat = t.reshape(-1)
    nz = torch.nonzero(flat).reshape(-1)[::stride]
    return nz.numpy().astype(np.int64), flat[nz].numpy()


def run_case(name):
    sub, w, fr, vol, t_rand, rk = build_case(name)
    net = ref_shim.build_reference_network(sub, w)
    net_mod, _cfg = ref_shim.load_reference()
    cap = {}

    smf = net_mod.Network._sample_motion_fields
    r2o = net_mod.Network._raw2outputs

    def smf_hook(**kw):
        r = smf(**kw)
        cap["x_skel"], cap["mask"], cap["pts"] = r["x_skel"].detach(), r["fg_likelihood_mask"].detach(), kw["pts"].detach()
        return r

    def r2o_hook(raw, raw_mask, z_vals, rays_d, bgcolor=None):
        r = r2o(raw, raw_mask, z_vals, rays_d, bgcolor)
        cap["raw"], cap["z"] = raw.detach(), z_vals.detach()
        cap["weights"], cap["term"] = r[2].detach(), r[4].detach()
        return r

    net._sample_motion_fields = smf_hook
    net._raw2outputs = r2o_hook
    vol_g = vol.clone().requires_grad_(True)
    out = ref_shim.reference_render_rays(net, fr, vol_g, t_rand=t_rand, **rk)
    counter_delta = (net.point_counter.data - 1.0).clone()
    g = {}
    if rk["training"]:
        scalar_loss(out).backward()
        m = net.cnl_mlp.module
        ge = m.encoder.embeddings.grad
        gi, gv = _sparse(ge, stride=61)
        offs = w.offsets.tolist()
        g.update(g_emb_idx=gi, g_emb_val=gv,
                 g_emb_level_sum=np.array([ge[a:b].double().sum().item() for a, b in zip(offs[:-1], offs[1:])]),
                 g_emb_level_l2=np.array([ge[a:b].double().norm().item() for a, b in zip(offs[:-1], offs[1:])]))
        vi, vv = _sparse(vol_g.grad, stride=7)
        g.update(g_vol_idx=vi, g_vol_val=vv, g_vol_sum=np.float64(vol_g.grad.double().sum().item()),
                 g_vol_l2=np.float64(vol_g.grad.double().norm().item()))
        g["g_point_dist"] = net.point_dist.grad.numpy()
        for i, li in enumerate((0, 2, 4, 6)):
            g[f"g_pts_w{i}"] = m.pts_linears[li].weight.grad.numpy()
            g[f"g_pts_b{i}"] = m.pts_linears[li].bias.grad.numpy()
            g[f"g_rgb_w{i}"] = m.rgb_linears[li].weight.grad.numpy()
            g[f"g_rgb_b{i}"] = m.rgb_linears[li].bias.grad.numpy()
        g["g_geo_w"], g["g_geo_b"] = m.geo_linear[0].weight.grad.numpy(), m.geo_linear[0].bias.grad.numpy()
        g["g_out_w"], g["g_out_b"] = m.output_linear[0].weight.grad.numpy(), m.output_linear[0].bias.grad.numpy()
        # keep the file small: 256x256 layers as float16-free sub-blocks
        for i in (1, 2, 3):
            g[f"g_pts_w{i}"] = g[f"g_pts_w{i}"][::8, ::8].copy()
            g[f"g_rgb_w{i}"] = g[f"g_rgb_w{i}"][::8, ::8].copy()

    # voxel bins by the reference's own arithmetic (network.py:367-369 + ATen's align_corners=True
    # un-normalisation); the reference never materialises them, grid_sample floors internally.
    pts = cap["pts"].reshape(-1, 3)
    bins = []
    for i in range(24):
        pos = torch.matmul(fr.motion_scale_Rs[i], pts.T).T + fr.motion_Ts[i]
        pos = (pos - fr.cnl_bbox_min_xyz[None, :]) * fr.cnl_bbox_scale_xyz[None, :] - 1.0
        bins.append(torch.floor(((pos + 1) / 2) * 31).to(torch.int32))
    bins = torch.stack(bins, 1)

    data = dict(
        rays_o=fr.rays_o.numpy(), rays_d=fr.rays_d.numpy(), near=fr.near.numpy(), far=fr.far.numpy(),
        motion_scale_Rs=fr.motion_scale_Rs.numpy(), motion_Ts=fr.motion_Ts.numpy(),
        emb_checksum=np.float64(w.embeddings.double().sum().item()),
        vol_checksum=np.float64(vol.double().sum().item()),
        z=cap["z"].numpy(), x_skel=cap["x_skel"].numpy(), mask=cap["mask"].numpy()[..., 0], bins=bins.numpy().astype(np.int16),
        raw=cap["raw"].numpy(), weights=cap["weights"].numpy(), term=cap["term"].numpy()[:, 0].astype(np.int32),
        rgb=out["rgb"].detach().numpy(), alpha=out["alpha"].detach().numpy(), depth=out["depth"].detach().numpy(),
        comp_loss=out["comp_loss"].detach().numpy(), counter_delta=counter_delta.numpy(), **g)
    if t_rand is not None:
        data["t_rand"] = t_rand.numpy()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, f"render_{name}.npz")
    np.savez_compressed(path, **data)
    print(f"{name}: rays={fr.rays_o.shape[0]} alpha.max={float(out['alpha'].max()):.4f} depth.max={float(out['depth'].max()):.4f} "
          f"counter+={int(counter_delta.sum())} -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    warnings.filterwarnings("ignore")
    for n in (sys.argv[1:] or CASES):
        run_case(n)
