"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU (torch fp32) restatement of OccNeRF's per-ray rendering path, written from the reference's
algorithm, not its code.  Each function cites the reference lines it follows.  Gradients come from
torch autograd on CPU; the hash grid (which has no CPU implementation upstream) is
oracle/hashgrid_oracle.c behind `hashgrid_c.HashGridFn`.

Pinned (oracle/make_golden.py -> tests/golden/*.npz, tests/test_oracle_golden.py) against the
UNMODIFIED reference functions imported through oracle/ref_shim.py in the build container:
`_get_samples_along_ray`, `_stratified_sampling`, `_sample_motion_fields` (ATen grid_sample),
`_apply_mlp_kernals`, `CanonicalMLP.forward`, `NonRigidMotionMLP.forward`, `_raw2outputs` and the whole
`_render_rays` forward + backward.  Two pieces of third-party arithmetic are NOT under /root/reference
and stay "parity unpinned": pykeops' K-min reduction (requirements.txt:10, unpinned version; restated
here as exact brute force, distance (dx*dx+dy*dy)+dz*dz in fp32, ties to the lowest index) and the
device `exp2f` inside gridencoder.cu:138 (handled by passing the device's own 16-entry scale table).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from oracle import hashgrid_c

S_DEFAULT = 128


# ----------------------------------------------------------------------------- a2/a3/a4 sampling
def z_samples(near, far, S=S_DEFAULT, t_rand=None):
    """network.py:416-432.  z = near*(1-t) + far*t with t = linspace(0,1,S); optional stratified jitter
    z' = lower + (upper-lower)*u with mids between neighbours."""
    t = torch.linspace(0.0, 1.0, steps=S).to(near)
    z = near * (1.0 - t) + far * t
    if t_rand is not None:
        mid = 0.5 * (z[:, 1:] + z[:, :-1])
        upper = torch.cat([mid, z[:, -1:]], -1)
        lower = torch.cat([z[:, :1], mid], -1)
        z = lower + (upper - lower) * t_rand
    return z


def sample_points(rays_o, rays_d, z):
    """network.py:456: p = o + d*z (un-normalised d), separate multiply and add."""
    return rays_o[:, None, :] + rays_d[:, None, :] * z[:, :, None]


# ----------------------------------------------------------------------------- a5 inverse-LBS warp
def _affine_fma(R, T, p):
    """q = R.p + T with the reduction order the reference's sgemm uses for K=3
    (network.py:367; verified against MKL here): acc = r0*x; acc = fma(r1,y,acc); acc = fma(r2,z,acc);
    then a separate add of T.  fp64 emulates the single-rounded fp32 FMA."""
    Rd, pd = R.double(), p.double()
    acc = (Rd[:, None, :, 0] * pd[None, :, None, 0]).float()
    acc = (Rd[:, None, :, 1] * pd[None, :, None, 1] + acc.double()).float()
    acc = (Rd[:, None, :, 2] * pd[None, :, None, 2] + acc.double()).float()
    return acc + T[:, None, :]


def lbs_warp(pts, Rs, Ts, vol, bbox_min, bbox_scale, *, exact=True, return_bins=False):
    """network.py:351-402 (`_sample_motion_fields`) restated for all 24 bones at once.

    pts (M,3); Rs (24,3,3); Ts (24,3); vol (25,32,32,32) [bone][z][y][x]; returns x_skel (M,3), mask (M,)
    and optionally the integer voxel bins floor(ix,iy,iz) (M,24,3) int32.
    Trilinear sampling follows ATen grid_sampler_3d with align_corners=True, zeros padding:
    ix = ((g+1)/2)*(W-1); corner weights (x1-ix)(y1-iy)(z1-iz) ...; out-of-range corners contribute 0.
    """
    w24 = vol[:-1]
    nb, D, H, W = w24.shape
    if exact:
        q = _affine_fma(Rs, Ts, pts)                                   # (24,M,3)
    else:
        q = torch.matmul(Rs, pts.T).transpose(1, 2) + Ts[:, None, :]
    g = (q - bbox_min) * bbox_scale - 1.0
    ix = ((g[..., 0] + 1) / 2) * (W - 1)
    iy = ((g[..., 1] + 1) / 2) * (H - 1)
    iz = ((g[..., 2] + 1) / 2) * (D - 1)
    x0, y0, z0 = torch.floor(ix), torch.floor(iy), torch.floor(iz)
    fx1, fy1, fz1 = ix - x0, iy - y0, iz - z0                           # weight of the +1 corner
    fx0, fy0, fz0 = (x0 + 1) - ix, (y0 + 1) - iy, (z0 + 1) - iz         # weight of the floor corner
    x0l, y0l, z0l = x0.long(), y0.long(), z0.long()
    flat = w24.reshape(nb, -1)
    bone = torch.arange(nb)[:, None]
    wsum = torch.zeros_like(ix)
    for dz, wz in ((0, fz0), (1, fz1)):
        for dy, wy in ((0, fy0), (1, fy1)):
            for dx, wx in ((0, fx0), (1, fx1)):
                xi, yi, zi = x0l + dx, y0l + dy, z0l + dz
                ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H) & (zi >= 0) & (zi < D)
                lin = (zi.clamp(0, D - 1) * H + yi.clamp(0, H - 1)) * W + xi.clamp(0, W - 1)
                val = flat[bone, lin]
                wsum = wsum + torch.where(ok, val * (wx * wy * wz), torch.zeros_like(val))
    total = wsum.sum(0)                                                  # (M,)
    x_skel = (wsum[..., None] * q).sum(0) / total.clamp(min=1e-4)[:, None]
    if return_bins:
        bins = torch.stack([x0l, y0l, z0l], -1).permute(1, 0, 2).to(torch.int32).contiguous()
        return x_skel, total, bins
    return x_skel, total


# ----------------------------------------------------------------------------- a8 non-rigid offsets
def hann_pe(xyz, iter_val, kick_in_iter, full_band_iter, multires=6):
    """embedders/hannw_fourier.py:15-45: [w_j sin(2^j x), w_j cos(2^j x)]_j, Hann-windowed, no identity term."""
    freqs = 2.0 ** torch.linspace(0.0, multires - 1, steps=multires)
    t = torch.clamp(torch.as_tensor(float(iter_val)) - torch.tensor(float(kick_in_iter)), min=0.0)
    alpha = multires * t / (full_band_iter - torch.tensor(float(kick_in_iter)))
    feats = []
    for j in range(multires):
        w = (1.0 - torch.cos(np.pi * torch.clamp(alpha - j, min=0.0, max=1.0))) / 2.0
        feats += [w * torch.sin(xyz * freqs[j]), w * torch.cos(xyz * freqs[j])]
    return torch.cat(feats, -1)


def hann_weights(iter_val, kick_in_iter, full_band_iter, multires=6):
    t = max(float(iter_val) - float(kick_in_iter), 0.0)
    alpha = torch.tensor(multires * t, dtype=torch.float32) / torch.tensor(float(full_band_iter) - float(kick_in_iter))
    return torch.stack([(1.0 - torch.cos(np.pi * torch.clamp(alpha - j, min=0.0, max=1.0))) / 2.0 for j in range(multires)])


def non_rigid_offsets(xyz, cond, pe, nr_w, nr_b):
    """non_rigid_motion_mlps/mlp_offset.py:45-62: [cond69, pe36] -> 128 x4 -> cat pe -> 128 -> 128 -> 3."""
    h = torch.cat([cond.expand(xyz.shape[0], -1), pe], -1)
    for i in range(6):
        if i == 4:
            h = torch.cat([h, pe], -1)
        h = F.relu(F.linear(h, nr_w[i], nr_b[i]))
    return F.linear(h, nr_w[6], nr_b[6])


# ----------------------------------------------------------------------------- a9 KNN
def knn_bruteforce(q, s, k, chunk=2048, stable=True):
    """Exact k nearest supports per query, ascending, ties to the lowest support index.
    Distance (dx*dx + dy*dy) + dz*dz in fp32 (restates the Kmin_argKmin call of knn.py:53,83)."""
    q, s = q.detach().float(), s.detach().float()
    out = torch.empty(q.shape[0], k, dtype=torch.int64)
    for i in range(0, q.shape[0], chunk):
        c = q[i:i + chunk]
        dx = c[:, None, 0] - s[None, :, 0]
        dy = c[:, None, 1] - s[None, :, 1]
        dz = c[:, None, 2] - s[None, :, 2]
        d = (dx * dx + dy * dy) + dz * dz
        if stable:
            out[i:i + chunk] = torch.sort(d, dim=1, stable=True)[1][:, :k]
        else:
            out[i:i + chunk] = torch.topk(d, k, dim=1, largest=False, sorted=True)[1]
    return out


def multiscale_knn(xyz, point_base, fps_index, k=10, stable=True):
    """network.py:236-255: k-NN in the full vertex set and in three FPS subsets, all ids mapped back to
    full-resolution vertex ids -> (m,4,k) int64."""
    levels = [knn_bruteforce(xyz, point_base, k, stable=stable)]
    for f in fps_index:
        levels.append(f[knn_bruteforce(xyz, point_base[f], k, stable=stable)])
    return torch.stack(levels, 1)


# ----------------------------------------------------------------------------- a10 per-vertex block
def vertex_block(point_base, point_dist, point_norms, stable=True):
    """network.py:263-284: 3-NN of the learnable cloud in the base cloud, |cos|-weighted projection,
    inside vote (>1.5 of 3) and signed mean distance.  Returns point_cloud (V,3), knn_base (V,3), dist (V,1)."""
    pc = point_base + point_dist
    kidx = knn_bruteforce(pc, point_base, 3, stable=stable)
    b = point_base[kidx]                                  # (V,3,3)
    direction = pc[:, None, :] - b
    n = point_norms[kidx]
    a = torch.abs(F.cosine_similarity(direction, n, dim=-1))[..., None]
    knn_base = (a * b).sum(1) / a.sum(1)
    inside = ((direction * n).sum(-1) < 0).sum(1) > 1.5
    dist = direction.norm(dim=-1).mean(1, keepdim=True)
    dist = torch.where(inside[:, None], -dist, dist)
    return pc, knn_base, dist


# ----------------------------------------------------------------------------- a11/a12 canonical MLP
def sample_geometry(xyz, knn0, point_base, point_norms, bound):
    """canonical_mlps/occnerf_mlp.py:146-167.  knn0 (m,10) = level-0 neighbours.  Returns the 4-D hash-grid
    input [p(3), normed_dist(1)] and the signed distance `dist` (m,1).  Everything is non-differentiable."""
    with torch.no_grad():
        P = point_base[knn0]
        Nn = point_norms[knn0]
        direction = xyz[:, None, :] - P
        inside = ((direction.double() * Nn.double()).sum(-1) < 0).sum(1) > knn0.shape[1] * 0.5
        dist = direction.norm(dim=-1).mean(1, keepdim=True)
        dist = torch.where(inside[:, None], -dist, dist)
        nd = torch.clamp((dist + 0.2) / 0.5, 0.0, 1.0)
        Pn = (P + bound) / (2 * bound)
        a = torch.abs(F.cosine_similarity(direction[:, :3], Nn[:, :3], dim=-1))[..., None]
        p = (a * Pn[:, :3]).sum(1) / a.sum(1)
        return torch.cat([p, nd], -1).float(), dist


def visibility_attention(point_counter, knn_idxs):
    """occnerf_mlp.py:110-115 (`simple_agg`): shift so the minimum is 1, scale so the maximum is 1,
    unbiased variance over the 40 neighbours, softmax."""
    att = point_counter[knn_idxs].reshape(knn_idxs.shape[0], -1)
    att = att + (1.0 - att.min(1, keepdim=True)[0])
    att = att / att.max(1, keepdim=True)[0]
    var = att.var(1, keepdim=True)
    return torch.softmax(att, 1), var


def canonical_mlp(agg, var, h, w):
    """occnerf_mlp.py:183-199: geo trunk [agg35,var1,h32] -> 256x4 -> 65; rgb trunk [geo64,agg35,h32] -> 256x4 -> 3."""
    x = torch.cat([agg, var, h], -1)
    for W_, b_ in zip(w.pts_w, w.pts_b):
        x = F.relu(F.linear(x, W_, b_))
    g = F.linear(x, w.geo_w, w.geo_b)
    sigma = g[:, :1]
    x = torch.cat([g[:, 1:], agg, h], -1)
    for W_, b_ in zip(w.rgb_w, w.rgb_b):
        x = F.relu(F.linear(x, W_, b_))
    rgb = F.linear(x, w.out_w, w.out_b)
    return rgb, sigma


def hash_encode(x4, w, level_scales=None):
    S = float(np.log2(w.per_level_scale))
    return hashgrid_c.HashGridFn.apply(x4, w.embeddings, w.offsets, S, 16, level_scales)


def query_canonical(xyz, subject, w, *, stable=True, level_scales=None, vb=None, return_aux=False):
    """network.py:236-299 + occnerf_mlp.py:142-199 for one chunk of (already offset) canonical points."""
    knn_idxs = multiscale_knn(xyz, subject.point_base, subject.fps_index, 10, stable=stable)
    pc, knn_base, dist_v = vb if vb is not None else vertex_block(subject.point_base, subject.point_dist,
                                                                  subject.point_norms, stable=stable)
    enc_in, dist = sample_geometry(xyz.detach(), knn_idxs[:, 0], subject.point_base, subject.point_norms, subject.bound)
    h = hash_encode(enc_in, w, level_scales)
    v_in = torch.cat([(knn_base + subject.bound) / (2 * subject.bound),
                      torch.clamp((dist_v + 0.2) / 0.8, 0.0, 1.0)], -1).float()
    feats = torch.cat([hash_encode(v_in, w, level_scales), pc.float()], -1)        # (V,35)
    att, var = visibility_attention(subject.point_counter.detach(), knn_idxs)
    agg = (att.detach()[..., None] * feats[knn_idxs.reshape(knn_idxs.shape[0], -1)]).sum(1)
    rgb, sigma = canonical_mlp(agg, var, h, w)
    raw = torch.cat([rgb, sigma, dist.detach()], -1)
    if return_aux:
        return raw, dict(knn_idxs=knn_idxs, enc_in=enc_in, att=att, var=var, agg=agg, h=h, feats=feats)
    return raw


# ----------------------------------------------------------------------------- a13 compositing
def composite(raw, mask, z, rays_d, bgcolor):
    """network.py:320-348 (`_raw2outputs`)."""
    delta = torch.cat([z[:, 1:] - z[:, :-1], torch.full_like(z[:, :1], 1e10)], -1) * rays_d.norm(dim=-1, keepdim=True)
    alpha = (1.0 - torch.exp(-F.softplus(raw[..., 3]) * delta)) * mask
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1.0 - alpha + 1e-10], -1), -1)[:, :-1]
    wgt = alpha * trans
    acc = wgt.sum(-1)
    rgb = (wgt[..., None] * torch.sigmoid(raw[..., :3])).sum(-2) + (1.0 - acc[:, None]) * bgcolor[None, :] / 255.0
    depth = (wgt * z).sum(-1)
    term = torch.argmax(alpha, dim=1)
    return rgb, acc, depth, term, wgt


def completeness_term(raw):
    """network.py:486-499: 10 * [dist<0] * exp(clamp(-relu(sigma_pre * [dist<=0.3]), -10, 0))."""
    dist = raw[..., 4]
    sig = torch.where(dist > 0.3, torch.zeros_like(raw[..., 3]), raw[..., 3])
    return (dist < 0).float().detach() * torch.exp(torch.clamp(-F.relu(sig), min=-10, max=0)) * 10.0


def visibility_hits(depth, term, x_skel, point_cloud, k=10, stable=True):
    """network.py:502-517: for rays with depth > 0.5 (only if more than one such ray) the canonical sample at
    argmax(alpha) votes for its 10 nearest learnable points; duplicates collapse.  Returns a (V,) 0/1 mask."""
    hits = torch.zeros(point_cloud.shape[0])
    sel = depth.detach() > 0.5
    if int(sel.sum()) > 1:
        pts = x_skel[sel, :, :][torch.arange(int(sel.sum())), term[sel]]
        idx = knn_bruteforce(pts, point_cloud, k, stable=stable)
        hits[idx.reshape(-1)] = 1.0
    return hits


# ----------------------------------------------------------------------------- a15 whole path
def render_rays(frame, vol, subject, w, *, iter_val, training, t_rand=None, kick_in_iter=100000,
                full_band_iter=200000, ignore_non_rigid=False, exact=True, stable=True, level_scales=None,
                S=S_DEFAULT, chunk=300000, return_aux=False):
    """network.py:435-525 (`_render_rays`) + 164-304 for one batch of rays.  Returns the reference's dict plus
    `term` and `hits` (the point_counter increment the reference applies in place, network.py:517)."""
    z = z_samples(frame.near, frame.far, S, t_rand)
    pts = sample_points(frame.rays_o, frame.rays_d, z)
    N = pts.shape[0]
    x_skel, mask = lbs_warp(pts.reshape(-1, 3), frame.motion_scale_Rs, frame.motion_Ts, vol,
                            frame.cnl_bbox_min_xyz, frame.cnl_bbox_scale_xyz, exact=exact)
    xyz_all = x_skel
    cond = frame.dst_posevec[None] if iter_val >= kick_in_iter else torch.zeros(1, 69)
    vb = vertex_block(subject.point_base, subject.point_dist, subject.point_norms, stable=stable)
    raws = []
    for i in range(0, xyz_all.shape[0], chunk):
        xyz = xyz_all[i:i + chunk]
        if not ignore_non_rigid:
            pe = hann_pe(xyz, iter_val, kick_in_iter, full_band_iter)
            xyz = xyz + non_rigid_offsets(xyz, cond, pe, w.nr_w, w.nr_b)
        raws.append(query_canonical(xyz, subject, w, stable=stable, level_scales=level_scales, vb=vb))
    raw = torch.cat(raws, 0).reshape(N, S, 5)
    rgb, acc, depth, term, _ = composite(raw, mask.reshape(N, S), z, frame.rays_d, frame.bgcolor)
    out = {"rgb": rgb, "alpha": acc, "depth": depth, "term": term}
    if training:
        out["comp_loss"] = completeness_term(raw)
        out["hits"] = visibility_hits(depth, term, x_skel.reshape(N, S, 3).detach(), vb[0].detach(), stable=stable)
    else:
        out["comp_loss"] = torch.zeros(1, 1)
    if return_aux:
        out.update(z=z, x_skel=x_skel.reshape(N, S, 3), mask=mask.reshape(N, S), raw=raw)
    return out
