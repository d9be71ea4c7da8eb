"""Operator layer: one Python function per C-ABI entry point (tensors in, tensors out) plus the
autograd.Function for the hash grid.  Everything below runs on the CUDA library; nothing here has a CPU path.

Shapes follow the reference: N rays, S samples per ray, M = N*S samples,
V vertices, k = 10 neighbours on 4 levels.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np
import torch

from occnerf_b200 import _lib
from occnerf_b200._lib import call, ptr, stream

f32, i32, i64, u8 = torch.float32, torch.int32, torch.int64, torch.uint8

_T_LIN = {}


def t_lin(S: int, device) -> torch.Tensor:
    """linspace(0,1,S) as the host framework computes it on the CPU (network.py:417), cached on the device."""
    key = (S, str(device))
    if key not in _T_LIN:
        _T_LIN[key] = torch.linspace(0.0, 1.0, steps=S).to(device)
    return _T_LIN[key]


# ----------------------------------------------------------------------------- K1 warp
def warp_pack_volume(vol, nb):
    """[>=nb, D,H,W] reference-layout weight volume -> corner-packed vol8 [nb, D+1, H+1, W+1, 8] (csrc/warp.cu): once per frame."""
    vd, vh, vw = vol.shape[-3:]
    if vol.shape[0] < nb:
        raise RuntimeError(f"warp_pack_volume: the weight volume has {vol.shape[0]} channels for {nb} bones")
    vol8 = torch.empty(nb, vd + 1, vh + 1, vw + 1, 8, device=vol.device, dtype=f32)
    call("occnerf_warp_pack_volume", ptr(vol, f32), nb, vd, vh, vw, ptr(vol8), stream())
    return vol8


def warp_forward(rays, t_rand, Rs, Ts, vol, bbox_min, bbox_scale, S, want_bins=False, vol8=None):
    """rays (N,8) -> z (N,S), x_skel (N,S,3), mask (N,S) [, bins (N,S,nb,3) int32].
    Default: the corner-packed kernel (`vol8` = warp_pack_volume(vol), built here when not given).  With `want_bins` the
    scalar-gather kernel on the reference layout runs (it is the one that can dump the voxel bins; same outputs)."""
    N, nb = rays.shape[0], Rs.shape[0]
    dev = rays.device
    z = torch.empty(N, S, device=dev, dtype=f32)
    x_skel = torch.empty(N, S, 3, device=dev, dtype=f32)
    mask = torch.empty(N, S, device=dev, dtype=f32)
    bins = torch.empty(N, S, nb, 3, device=dev, dtype=i32) if want_bins else None
    vd, vh, vw = vol.shape[-3:]
    if vol.shape[0] < nb:
        raise RuntimeError(f"warp_forward: the weight volume has {vol.shape[0]} channels for {nb} bones")
    if not want_bins:
        if vol8 is None:
            vol8 = warp_pack_volume(vol, nb)
        call("occnerf_warp_forward_packed", ptr(rays, f32), ptr(t_lin(S, dev), f32), ptr(t_rand, f32), ptr(Rs, f32), ptr(Ts, f32),
             ptr(vol8, f32), ptr(bbox_min, f32), ptr(bbox_scale, f32), N, S, nb, vd, vh, vw, ptr(z), ptr(x_skel), ptr(mask), stream())
        return z, x_skel, mask
    call("occnerf_warp_forward", ptr(rays, f32), ptr(t_lin(S, dev), f32), ptr(t_rand, f32), ptr(Rs, f32), ptr(Ts, f32),
         ptr(vol, f32), ptr(bbox_min, f32), ptr(bbox_scale, f32), N, S, nb, vd, vh, vw, ptr(z), ptr(x_skel), ptr(mask),
         ptr(bins), stream())
    return (z, x_skel, mask, bins) if want_bins else (z, x_skel, mask)


def warp_backward(rays, t_rand, Rs, Ts, bbox_min, bbox_scale, g_mask, S, vol_shape, vol8=None, want_pose=False, packed=True):
    """d mask (N,S) -> g_vol with the shape of the reference's motion_weights_vol (channels >= nb stay 0)
    [, g_Rs (nb,3,3), g_Ts (nb,3) with want_pose: the gradient F.grid_sample gives the reference w.r.t. its grid, chained
    through q = R p + T].  packed=False runs the scalar-reduction kernel on the reference layout (cross-check)."""
    N, nb = rays.shape[0], Rs.shape[0]
    vd, vh, vw = vol_shape[-3:]
    dev = rays.device
    if packed:
        if want_pose and vol8 is None:
            raise RuntimeError("warp_backward: the pose gradients need the packed volume (vol8)")
        g_vol8 = torch.zeros(nb, vd + 1, vh + 1, vw + 1, 8, device=dev, dtype=f32)
        g_Rs = torch.zeros(nb, 3, 3, device=dev, dtype=f32) if want_pose else None
        g_Ts = torch.zeros(nb, 3, device=dev, dtype=f32) if want_pose else None
        call("occnerf_warp_backward_packed", ptr(rays, f32), ptr(t_lin(S, dev), f32), ptr(t_rand, f32), ptr(Rs, f32), ptr(Ts, f32),
             ptr(vol8, f32), ptr(bbox_min, f32), ptr(bbox_scale, f32), ptr(g_mask, f32), N, S, nb, vd, vh, vw, ptr(g_vol8), ptr(g_Rs),
             ptr(g_Ts), stream())
        g_vol = torch.empty(vol_shape, device=dev, dtype=f32)
        call("occnerf_warp_unpack_grad", ptr(g_vol8), nb, int(vol_shape[0]), vd, vh, vw, ptr(g_vol), stream())
        return (g_vol, g_Rs, g_Ts) if want_pose else g_vol
    g_vol = torch.zeros(vol_shape, device=rays.device, dtype=f32)
    call("occnerf_warp_backward", ptr(rays, f32), ptr(t_lin(S, rays.device), f32), ptr(t_rand, f32), ptr(Rs, f32),
         ptr(Ts, f32), ptr(bbox_min, f32), ptr(bbox_scale, f32), ptr(g_mask, f32), N, S, nb, vd, vh, vw, ptr(g_vol),
         stream())
    return g_vol


# ----------------------------------------------------------------------------- per-frame prologue (csrc/prologue.cu)
def motion_basis(dst_Rs, dst_Ts, cnl_gtfms):
    """(24,3,3), (24,3), (24,4,4) -> motion_scale_Rs (24,3,3), motion_Ts (24,3)  (network_util.py:138-200)."""
    nb = dst_Rs.shape[0]
    Rs = torch.empty(nb, 3, 3, device=dst_Rs.device, dtype=f32)
    Ts = torch.empty(nb, 3, device=dst_Rs.device, dtype=f32)
    call("occnerf_motion_basis", ptr(dst_Rs, f32), ptr(dst_Ts, f32), ptr(cnl_gtfms, f32), nb, ptr(Rs), ptr(Ts), stream())
    return Rs, Ts


def pose_refine(weights5, biases5, posevec69, dst_Rs):
    """BodyPoseRefiner + Rodrigues + dst_Rs[1:] . R_delta in one launch (forward only)."""
    wp = (C.c_void_p * 5)(*[ptr(w.detach().contiguous(), f32) for w in weights5])
    bp = (C.c_void_p * 5)(*[ptr(b.detach().contiguous(), f32) for b in biases5])
    out = torch.empty_like(dst_Rs)
    call("occnerf_pose_refine", C.cast(wp, C.c_void_p), C.cast(bp, C.c_void_p), ptr(posevec69, f32), ptr(dst_Rs, f32), dst_Rs.shape[0],
         ptr(out), stream())
    return out


def _deconv_splits(tiles: int, k_steps: int) -> int:
    """Split-K factor that brings a GEMM's grid to about two CTAs per SM (148 SMs), keeping at least 8 K-steps per split."""
    return max(1, min((296 + tiles - 1) // tiles, k_steps // 8))


def deconv3d_forward(W, bias, Yin, D, slope, exact=False):
    """Yout [Cout, (2D)^3] = bias + ConvTranspose3d(4, 2, 1)(LeakyReLU_slope(Yin)); W [Cin, Cout, 4, 4, 4], Yin [Cin, D^3] (batch 1)."""
    Cin, Cout = W.shape[0], W.shape[1]
    V, Q = D ** 3, Cout * (8 if D == 1 else 64)
    Yout = torch.empty(Cout, 8 * V, device=W.device, dtype=f32)
    tiles = ((Q + 127) // 128) if V <= 8 else ((Q + 63) // 64) * ((V + 63) // 64)
    call("occnerf_deconv3d_forward", ptr(W, f32), ptr(bias, f32), ptr(Yin, f32), Cin, Cout, D, float(slope), _deconv_splits(tiles, Cin // 16),
         int(exact), ptr(Yout), stream())
    return Yout


_deconv_overlap_set = False


def deconv3d_backward(W, Yin, dYout, D, slope, exact=False, need_dyin=True, dW_out=None):
    """-> (dW [Cin, Cout, 4, 4, 4], dbias [Cout], dYin [Cin, D^3] | None).  dW_out: caller's destination for dW (e.g. a view of the
    data-parallel all-reduce buffer); it is overwritten."""
    global _deconv_overlap_set
    if not _deconv_overlap_set:               # OCCNERF_DECONV_OVERLAP=0: no side stream (for A/B timing)
        _deconv_overlap_set = True
        if os.environ.get("OCCNERF_DECONV_OVERLAP", "1") == "0":
            call("occnerf_deconv_set_overlap", 0)
    Cin, Cout = W.shape[0], W.shape[1]
    V, Q = D ** 3, Cout * (8 if D == 1 else 64)
    bm = 256 if (Cin >= 256 and not exact) else 64     # row tile of the gradient GEMMs (csrc/deconv.cu launch_gemm)
    w_splits = _deconv_splits(((Cin + bm - 1) // bm) * ((Q + 63) // 64), max(V // 16, 1))
    d_tiles = ((Cin + 127) // 128) if V <= 8 else ((Cin + bm - 1) // bm) * ((V + 63) // 64)
    d_splits = _deconv_splits(d_tiles, Q // 16)
    need_zero = w_splits > 1 or D == 1        # D = 1: only the 8 reachable taps per (ci, o) are written, the other 56 are zero
    if dW_out is None:
        dW = torch.zeros_like(W) if need_zero else torch.empty_like(W)
    else:
        dW = dW_out
        if need_zero:
            dW.zero_()
    db = torch.empty(Cout, device=W.device, dtype=f32)
    dYin = None
    if need_dyin:
        dYin = torch.zeros(Cin, V, device=W.device, dtype=f32) if d_splits > 1 else torch.empty(Cin, V, device=W.device, dtype=f32)
    call("occnerf_deconv3d_backward", ptr(W, f32), ptr(Yin, f32), ptr(dYout, f32), Cin, Cout, D, float(slope), w_splits, d_splits, 0, int(exact),
         ptr(dW), ptr(db), ptr(dYin) if need_dyin else None, stream())
    return dW, db, dYin


def decoder_linear_forward(w, b, e):
    y = torch.empty(w.shape[0], device=w.device, dtype=f32)
    call("occnerf_decoder_linear_forward", ptr(w, f32), ptr(b, f32), ptr(e, f32), w.shape[0], w.shape[1], ptr(y), stream())
    return y


def decoder_linear_backward(w, e, g):
    dw, db, de = torch.empty_like(w), torch.empty(w.shape[0], device=w.device, dtype=f32), torch.empty(w.shape[1], device=w.device, dtype=f32)
    call("occnerf_decoder_linear_backward", ptr(w, f32), ptr(e, f32), ptr(g, f32), w.shape[0], w.shape[1], ptr(dw), ptr(db), ptr(de), stream())
    return dw, db, de


def weight_volume_forward(logits, priors):
    """softmax over the channels of logits + log(priors), [channels, D, H, W]."""
    vol = torch.empty_like(logits)
    call("occnerf_weight_volume_forward", ptr(logits, f32), ptr(priors, f32), logits.shape[0], logits[0].numel(), ptr(vol), stream())
    return vol


def weight_volume_backward(vol, g_vol):
    g = torch.empty_like(vol)
    call("occnerf_weight_volume_backward", ptr(vol, f32), ptr(g_vol, f32), vol.shape[0], vol[0].numel(), ptr(g), stream())
    return g


# ----------------------------------------------------------------------------- KNN
def to_float4(points: torch.Tensor) -> torch.Tensor:
    out = torch.zeros(points.shape[0], 4, device=points.device, dtype=f32)
    out[:, :3] = points
    return out


def knn(queries, supports4, level_begin, k, support_gid=None, query_sel=None, out=None):
    """queries (m,3) -> (m, n_levels, k) int32 neighbour ids (see occnerf_knn in include/occnerf_b200.h)."""
    m, nl = queries.shape[0], len(level_begin) - 1
    if out is None:
        out = torch.empty(m, nl, k, device=queries.device, dtype=i32)
    lb = (C.c_int32 * (nl + 1))(*[int(v) for v in level_begin])
    call("occnerf_knn", ptr(queries, f32), m, ptr(supports4, f32), ptr(support_gid, i32), C.cast(lb, C.c_void_p), nl, k,
         ptr(query_sel, u8), ptr(out, i32), stream())
    return out


def build_knn_hierarchy(fine: torch.Tensor, centers: torch.Tensor):
    """Static acceleration structure for occnerf_knn_hier: assigns every fine point to its nearest centre and returns
    (fine4 cluster-sorted with the original row in .w, centers4 with the inflated cluster radius in .w, ranges [nc,2])."""
    nf, nc, dev = fine.shape[0], centers.shape[0], fine.device
    diff = fine[:, None, :].float() - centers[None, :, :].float()
    d = torch.sqrt((diff * diff).sum(-1))                              # (nf, nc), explicit differences (no matmul trick)
    dmin, assign = d.min(1)
    order = torch.argsort(assign, stable=True)
    counts = torch.bincount(assign, minlength=nc)
    begin = torch.cumsum(counts, 0) - counts
    radius = torch.zeros(nc, device=dev, dtype=f32).scatter_reduce(0, assign, dmin, reduce="amax", include_self=True)
    radius = radius * 1.00001 + 1e-6
    fine4 = torch.empty(nf, 4, device=dev, dtype=f32)
    fine4[:, :3] = fine[order].float()
    fine4[:, 3] = order.to(i32).view(f32)
    centers4 = torch.cat([centers.float(), radius[:, None]], 1).contiguous()
    ranges = torch.stack([begin, counts], 1).to(i32).contiguous()
    return fine4.contiguous(), centers4, ranges


def knn_hier(queries, group_stride, fine4, centers4, ranges, out, fine_level, center_level, fine_gid=None, center_gid=None):
    """Writes the k-NN in the fine and the centre level into out[:, fine_level, :] and out[:, center_level, :]."""
    m, nl, k = out.shape
    base = out.data_ptr()
    call("occnerf_knn_hier", ptr(queries, f32), m, int(group_stride), ptr(fine4, f32), ptr(centers4, f32), ptr(ranges, i32),
         fine4.shape[0], centers4.shape[0], ptr(fine_gid, i32), ptr(center_gid, i32), k, base + 4 * k * fine_level,
         base + 4 * k * center_level, nl * k, stream())
    return out


def _nearest(fine, centers):
    diff = fine[:, None, :].float() - centers[None, :, :].float()
    return torch.sqrt((diff * diff).sum(-1)).min(1)                    # explicit differences (no matmul trick)


def _inflate(r):
    return r * 1.00001 + 1e-6


def build_knn_tree(base: torch.Tensor, fps):
    """Static three-level cluster tree for occnerf_knn_tree (see include/occnerf_b200.h for the table layout)."""
    dev = base.device
    P1, P2, P3 = base[fps[0]], base[fps[1]], base[fps[2]]
    n0, n1, n2, n3 = base.shape[0], P1.shape[0], P2.shape[0], P3.shape[0]

    def bits(t):
        return t.to(i32).view(f32)

    def amax(n, idx, val):
        return torch.zeros(n, device=dev, dtype=f32).scatter_reduce(0, idx, val, reduce="amax", include_self=True)

    d23, a23 = _nearest(P2, P3)
    order2 = torch.argsort(a23, stable=True)
    cnt32 = torch.bincount(a23, minlength=n3)
    beg32 = torch.cumsum(cnt32, 0) - cnt32
    inv2 = torch.empty(n2, device=dev, dtype=torch.int64)
    inv2[order2] = torch.arange(n2, device=dev)
    d13, a13 = _nearest(P1, P3)
    order1 = torch.argsort(a13, stable=True)
    cnt31 = torch.bincount(a13, minlength=n3)
    beg31 = torch.cumsum(cnt31, 0) - cnt31
    d02, a02 = _nearest(base, P2)
    pos0 = inv2[a02]                                                   # cluster of every vertex, as a position in p2s
    order0 = torch.argsort(pos0, stable=True)
    cnt20 = torch.bincount(pos0, minlength=n2)
    beg20 = torch.cumsum(cnt20, 0) - cnt20
    r20 = _inflate(amax(n2, pos0, d02))
    R30 = _inflate(amax(n3, a23[order2], d23[order2] + r20))

    def f4(pts, order):
        out = torch.empty(pts.shape[0], 4, device=dev, dtype=f32)
        out[:, :3] = pts[order].float()
        out[:, 3] = bits(order)
        return out.contiguous()

    zero2, zero3 = torch.zeros(n2, device=dev), torch.zeros(n3, device=dev)
    return dict(
        p0s=f4(base, order0), p1s=f4(P1, order1), p2s=f4(P2, order2), p3=to_float4(P3),
        c2tab=torch.stack([r20, bits(beg20), bits(cnt20), zero2], 1).contiguous(),
        c3tab=torch.stack([_inflate(amax(n3, a23, d23)), _inflate(amax(n3, a13, d13)), R30, zero3], 1).contiguous(),
        c3rng=torch.stack([beg32, cnt32, beg31, cnt31], 1).to(i32).contiguous(),
        gid1=fps[0].to(i32).contiguous(), gid2=fps[1].to(i32).contiguous(), gid3=fps[2].to(i32).contiguous(),
        inv2=inv2.to(i32).contiguous(), n=(n0, n1, n2, n3))


KNN_LANE_RAYS = 32        # cluster-tree kernels: 32 rays at one depth per warp
# grid kernel: a warp covers 2 rays x 16 consecutive depths (neighbouring depths share or neighbour the 12.5 mm candidate cell); a scheduling hint
# only, results do not depend on it.  786 k queries on B200: 2 -> 0.343 ms, 4 -> 0.351, 8 -> 0.364, 16 -> 0.376, 32 -> 0.426
KNN_GRID_LANE_RAYS = int(os.environ.get("OCCNERF_KNN_LANE_RAYS", "2"))


def knn_tree(queries, group_stride, tree, out=None, lane_rays=None):
    """All 4 levels x k=10 in one launch -> (m,4,10) int32 vertex ids."""
    lane_rays = KNN_LANE_RAYS if lane_rays is None else lane_rays
    m = queries.shape[0]
    if out is None:
        out = torch.empty(m, 4, 10, device=queries.device, dtype=i32)
    t = tree
    call("occnerf_knn_tree", ptr(queries, f32), m, int(group_stride), int(lane_rays), ptr(t["p0s"], f32), ptr(t["p1s"], f32), ptr(t["p2s"], f32),
         ptr(t["p3"], f32), ptr(t["c2tab"], f32), ptr(t["c3tab"], f32), ptr(t["c3rng"], i32), ptr(t["gid1"], i32),
         ptr(t["gid2"], i32), ptr(t["gid3"], i32), ptr(t["inv2"], i32), *t["n"], 10, ptr(out, i32), stream())
    return out


_GRID_CACHE = {}


# Edge of a candidate-grid cell (m).  Smaller cells = shorter candidate lists (the kernel's work) for more list memory: B200, 786 k
# queries: 0.025 -> 0.50 ms, 0.0175 -> 0.41, 0.0125 -> 0.36, 0.01 -> 0.34, 0.008 -> 0.32 (gpurun_out/bench_cell_*.json; the build of
# the two smallest takes tens of seconds).
KNN_GRID_CELL = float(os.environ.get("OCCNERF_KNN_CELL", "0.0125"))


def build_knn_grid(base: torch.Tensor, fps, cell: float = None, pad: float = 0.35, k: int = 10):
    """Static candidate lists for occnerf_knn_grid (include/occnerf_b200.h).  A uniform grid of `cell`-sized cells covers the
    vertex bounding box +- `pad`; for every cell (centre o, half diagonal rho) and level the list holds every point p with
        |p - o| <= d_k(o) + 2*rho + margin,     d_k(o) = distance from o to its k-th nearest point of the level.
    Superset proof: a query q in the cell has |q - o| <= rho, hence d_k(q) <= d_k(o) + rho (the k points within d_k(o) of
    o are within d_k(o) + rho of q), and every k-nearest p of q satisfies |p - o| <= |p - q| + rho <= d_k(o) + 2*rho.  The
    margin (1e-4 + 1e-5*d) is far above fp32 rounding.  Lists are sorted by |p - o| and padded to nothing; built once per
    subject on the device (a few seconds, ~0.3 GB at the defaults)."""
    dev = base.device
    cell = KNN_GRID_CELL if cell is None else cell
    key = (str(dev), tuple(base.shape), float(base.double().sum()), float(base.double().abs().sum()), cell, pad, k,
           tuple(int(f.shape[0]) for f in fps), tuple(int(f.long().sum()) for f in fps))
    if key in _GRID_CACHE:
        return _GRID_CACHE[key]
    base = base.float()
    levels = [base, base[fps[0]], base[fps[1]], base[fps[2]]]
    gmin = base.min(0)[0] - pad
    dims = torch.ceil((base.max(0)[0] + pad - gmin) / cell).long().tolist()
    ncell = dims[0] * dims[1] * dims[2]
    rho = cell * math.sqrt(3.0) / 2.0
    idx = torch.arange(ncell, device=dev)
    ix, iy, iz = idx % dims[0], (idx // dims[0]) % dims[1], idx // (dims[0] * dims[1])
    centres = torch.stack([ix, iy, iz], 1).float().add_(0.5).mul_(cell).add_(gmin)
    cell_tab = torch.zeros(ncell, 4, 2, device=dev, dtype=i32)
    lists, total = [], 0
    for lev in range(4):
        P = levels[lev]
        n = P.shape[0]
        chunk = max(256, (64 << 20) // n)                      # ~256 MB of fp32 distances per chunk
        for c0 in range(0, ncell, chunk):
            c = centres[c0:c0 + chunk]
            d = torch.zeros(c.shape[0], n, device=dev)
            for a in range(3):
                d += (c[:, a:a + 1] - P[None, :, a]) ** 2
            d.sqrt_()
            ds, order = torch.sort(d, dim=1)
            thr = ds[:, min(k, n) - 1]
            thr = thr + 2.0 * rho + 1e-4 + 1e-5 * thr
            cnt = (ds <= thr[:, None]).sum(1)
            width = int(cnt.max())
            keep = torch.arange(width, device=dev)[None, :] < cnt[:, None]
            lists.append(order[:, :width][keep].to(torch.int16))
            off = torch.cumsum(cnt, 0) - cnt + total
            cell_tab[c0:c0 + chunk, lev, 0] = off.to(i32)
            cell_tab[c0:c0 + chunk, lev, 1] = cnt.to(i32)
            total += int(cnt.sum())
            del d, ds, order, keep
    if total >= 2 ** 31:
        raise RuntimeError(f"build_knn_grid: {total} candidate entries overflow the int32 offsets; use a coarser cell")
    grid = dict(
        p=[to_float4(P).contiguous() for P in levels], n=tuple(int(P.shape[0]) for P in levels),
        gid=[f.to(i32).contiguous() for f in fps], cell_tab=cell_tab.contiguous(), lists=torch.cat(lists).contiguous(),
        params=(C.c_float * 4)(float(gmin[0]), float(gmin[1]), float(gmin[2]), float(np.float32(1.0) / np.float32(cell))),
        dims=(C.c_int32 * 3)(*dims), entries=total, cells=ncell)
    _GRID_CACHE[key] = grid
    return grid


def knn_grid_lane_rays(lane_rays: int, group_stride: int) -> int:
    """Rays per warp actually used: with fewer samples per ray than 32 / lane_rays the depth axis cannot fill the warp, so it takes more rays
    (group_stride = 1, flat query lists: 32 rays x 1 sample).  Powers of two <= 32, as occnerf_knn_grid requires."""
    lane_samples = max(1, 32 // max(1, lane_rays))
    while lane_samples > max(1, group_stride):
        lane_samples //= 2
    return 32 // lane_samples


def knn_grid(queries, group_stride, grid, out=None, lane_rays=None):
    """All 4 levels x k=10 through the per-cell candidate lists -> (m,4,10) int32 vertex ids (bit-identical to knn)."""
    m = queries.shape[0]
    lane_rays = KNN_GRID_LANE_RAYS if lane_rays is None else lane_rays
    lane_rays = knn_grid_lane_rays(lane_rays, int(group_stride))
    if out is None:
        out = torch.empty(m, 4, 10, device=queries.device, dtype=i32)
    g = grid
    call("occnerf_knn_grid", ptr(queries, f32), m, int(group_stride), int(lane_rays), ptr(g["p"][0], f32), ptr(g["p"][1], f32),
         ptr(g["p"][2], f32), ptr(g["p"][3], f32), *g["n"], ptr(g["gid"][0], i32), ptr(g["gid"][1], i32), ptr(g["gid"][2], i32),
         ptr(g["cell_tab"], i32), ptr(g["lists"], torch.int16), C.cast(g["params"], C.c_void_p), C.cast(g["dims"], C.c_void_p), 10,
         ptr(out, i32), stream())
    return out


def sample_geometry(xyz, knn_idx, point_base, point_norms, bound, raw=None):
    """-> enc_in (m,4), dist.  With `raw` (m,5) given, dist is written into raw[:,4] in place and returned as a view."""
    m = xyz.shape[0]
    enc_in = torch.empty(m, 4, device=xyz.device, dtype=f32)
    stride_knn = knn_idx.shape[1] * knn_idx.shape[2] if knn_idx.dim() == 3 else knn_idx.shape[1]
    if raw is None:
        dist = torch.empty(m, device=xyz.device, dtype=f32)
        dptr, dstride = ptr(dist), 1
    else:
        if tuple(raw.shape) != (m, 5) or not raw.is_contiguous():
            raise RuntimeError(f"sample_geometry: raw must be a contiguous ({m}, 5) tensor, got {tuple(raw.shape)}")
        dist = raw[:, 4]
        dptr, dstride = raw.data_ptr() + 16, 5
    call("occnerf_sample_geometry", ptr(xyz, f32), ptr(knn_idx, i32), stride_knn, ptr(point_base, f32),
         ptr(point_norms, f32), float(bound), m, ptr(enc_in), dptr, dstride, stream())
    return enc_in, dist


def vertex_block_forward(point_base, point_dist, point_norms, kidx3, bound, v_in, tail_ptr, ld):
    V = point_base.shape[0]
    call("occnerf_vertex_block_forward", ptr(point_base, f32), ptr(point_dist, f32), ptr(point_norms, f32), ptr(kidx3, i32),
         float(bound), V, ptr(v_in, f32), tail_ptr, ld, stream())


def vertex_block_backward(point_base, point_dist, point_norms, kidx3, bound, g_v_in, g_tail_ptr, ld):
    V = point_base.shape[0]
    g_pd = torch.empty(V, device=point_base.device, dtype=f32)
    call("occnerf_vertex_block_backward", ptr(point_base, f32), ptr(point_dist, f32), ptr(point_norms, f32), ptr(kidx3, i32),
         float(bound), V, ptr(g_v_in, f32), g_tail_ptr, ld, ptr(g_pd), stream())
    return g_pd


# ----------------------------------------------------------------------------- hash grid
_SCALES = {}


def level_scales(S: float, H: int, L: int, device) -> torch.Tensor:
    """Per-level scale table evaluated on the device with the reference's expression (gridencoder.cu:138)."""
    key = (float(np.float32(S)), int(H), int(L), str(device))
    if key not in _SCALES:
        out = torch.empty(L, device=device, dtype=f32)
        call("occnerf_hashgrid_level_scales", float(np.float32(S)), H, L, ptr(out), stream())
        _SCALES[key] = out
    return _SCALES[key]


def hashgrid_forward(inputs, embeddings, offsets, scales, *, out=None, out_ptr=None, ld=None, layout=_lib.LAYOUT_BLC,
                     want_dy_dx=False, want_cells=False, run_length=0):
    B, D = inputs.shape
    Cc, L = embeddings.shape[1], offsets.shape[0] - 1
    dev = inputs.device
    if out is None and out_ptr is None:
        out = torch.empty((L, B, Cc) if layout == _lib.LAYOUT_LBC else (B, L * Cc), device=dev, dtype=f32)
    if out_ptr is None:
        out_ptr = ptr(out, f32)
    if ld is None:
        ld = L * Cc
    dy_dx = torch.empty(B, L * D * Cc, device=dev, dtype=f32) if want_dy_dx else None
    cells = torch.empty(B, L, D, device=dev, dtype=i32) if want_cells else None
    slots = torch.empty(B, L, 1 << D, device=dev, dtype=i32) if want_cells else None
    call("occnerf_hashgrid_forward", ptr(inputs, f32), ptr(embeddings, f32), ptr(offsets, i32), ptr(scales, f32), out_ptr,
         layout, ld, B, D, Cc, L, ptr(dy_dx), ptr(cells), ptr(slots), int(run_length), stream())
    return out, dy_dx, cells, slots


HASH_BWD_RUN = 16


def hashgrid_backward(grad_ptr, ld, layout, inputs, offsets, scales, g_emb, Cc, run_length=0):
    """run_length 8/16: inputs are ordered along rays -> merge the reductions of consecutive samples in one cell."""
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    call("occnerf_hashgrid_backward", grad_ptr, layout, ld, ptr(inputs, f32), ptr(offsets, i32), ptr(scales, f32),
         ptr(g_emb, f32), B, D, Cc, L, int(run_length), stream())
    return g_emb


def hashgrid_input_backward(grad_ptr, ld, layout, dy_dx, B, D, Cc, L):
    g_in = torch.empty(B, D, device=dy_dx.device, dtype=f32)
    call("occnerf_hashgrid_input_backward", grad_ptr, layout, ld, ptr(dy_dx, f32), ptr(g_in), B, D, Cc, L, stream())
    return g_in


class _GridEncode(torch.autograd.Function):
    """Differentiable hash-grid encode with the reference's semantics (gridencoder/grid.py:24-90):
    inputs (B,D) in [0,1] -> (B, L*C); gradients to the table always, to the inputs iff they require grad."""

    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution):
        inputs = inputs.contiguous().float()
        S = float(np.log2(per_level_scale))
        scales = level_scales(S, base_resolution, offsets.shape[0] - 1, inputs.device)
        need_in = bool(ctx.needs_input_grad[0])       # (of the caller's tensor: the contiguous fp32 copy above never requires grad)
        out, dy_dx, _, _ = hashgrid_forward(inputs.detach(), embeddings.detach().contiguous(), offsets, scales,
                                            want_dy_dx=need_in)
        ctx.save_for_backward(inputs.detach(), offsets, scales, dy_dx if need_in else torch.empty(0, device=inputs.device))
        ctx.meta = (tuple(embeddings.shape), need_in)
        return out

    @staticmethod
    def backward(ctx, g):
        inputs, offsets, scales, dy_dx = ctx.saved_tensors
        (n_emb, Cc), need_in = ctx.meta
        g = g.contiguous()
        B, D = inputs.shape
        L = offsets.shape[0] - 1
        g_emb = torch.zeros(n_emb, Cc, device=g.device, dtype=f32)
        hashgrid_backward(ptr(g, f32), L * Cc, _lib.LAYOUT_BLC, inputs, offsets, scales, g_emb, Cc)
        g_in = hashgrid_input_backward(ptr(g, f32), L * Cc, _lib.LAYOUT_BLC, dy_dx, B, D, Cc, L) if need_in else None
        return g_in, g_emb, None, None, None


grid_encode = _GridEncode.apply


# ----------------------------------------------------------------------------- aggregation
def aggregate_forward(knn_idx, point_counter, feats36, X_ptr, ldx, want_att=False):
    """want_att=True also returns the attention weights (m,nn), which select the run-length backward (samples ordered
    along rays)."""
    m = knn_idx.shape[0]
    nn = knn_idx.numel() // max(m, 1)
    att_w = torch.empty(m, nn, device=knn_idx.device, dtype=f32) if want_att else None
    call("occnerf_aggregate_forward", ptr(knn_idx, i32), ptr(point_counter, f32), ptr(feats36, f32), m, nn, X_ptr, ldx,
         ptr(att_w), stream())
    return att_w


AGG_BWD_COPIES = 64


def aggregate_backward(knn_idx, point_counter, gX_ptr, ldg, V, copies=None, g_priv=None, att_w=None):
    """g_feats (V,36): one vector reduction per (sample, neighbour, column chunk) into `copies` privatised replicas.
    With `g_priv` (copies,V,36) given, accumulates into it and returns it (the caller sums the replicas)."""
    m = knn_idx.shape[0]
    nn = knn_idx.numel() // max(m, 1)
    if g_priv is not None:
        call("occnerf_aggregate_backward", ptr(knn_idx, i32), ptr(point_counter, f32), gX_ptr, ldg, m, nn, ptr(g_priv, f32), V,
             g_priv.shape[0], ptr(att_w), stream())
        return g_priv
    copies = AGG_BWD_COPIES if copies is None else copies
    g_priv = torch.zeros(copies, V, 36, device=knn_idx.device, dtype=f32)
    call("occnerf_aggregate_backward", ptr(knn_idx, i32), ptr(point_counter, f32), gX_ptr, ldg, m, nn, ptr(g_priv), V, copies,
         ptr(att_w), stream())
    return g_priv.sum(0) if copies > 1 else g_priv[0]


# ----------------------------------------------------------------------------- fp32 GEMM helpers
def sgemm(A_ptr, sAi, sAr, B_ptr, sBr, sBj, C_ptr, ldc, Mi, Nj, Kr, *, bias=None, mask_ptr=None, ldmask=0, relu=False,
          accum=False, split_k=1):
    flags = (_lib.GEMM_BIAS if bias is not None else 0) | (_lib.GEMM_RELU if relu else 0) | \
            (_lib.GEMM_ACCUM if accum else 0) | (_lib.GEMM_RELUMASK if mask_ptr is not None else 0)
    call("occnerf_sgemm", A_ptr, sAi, sAr, B_ptr, sBr, sBj, C_ptr, ldc, ptr(bias, f32) if bias is not None else None,
         mask_ptr, ldmask, Mi, Nj, Kr, flags, split_k, stream())


def colsum(A_ptr, lda, Mi, Nj, out, mask_ptr=None, ldmask=0):
    call("occnerf_colsum", A_ptr, lda, mask_ptr, ldmask, Mi, Nj, ptr(out, f32), stream())


def hann_window(iter_val, kick_in_iter, full_band_iter, multires=6):
    """Window weights of hannw_fourier.py:27-33 evaluated in fp32 like the reference's tensor arithmetic."""
    t = torch.clamp(torch.tensor(float(iter_val)) - torch.tensor(float(kick_in_iter)), min=0.0)
    alpha = multires * t / (full_band_iter - torch.tensor(float(kick_in_iter)))
    w = [(1.0 - torch.cos(math.pi * torch.clamp(alpha - j, min=0.0, max=1.0))) / 2.0 for j in range(multires)]
    return [float(v) for v in w]


def hann_pe(xyz, window, out=None):
    m, mr = xyz.shape[0], len(window)
    if out is None:
        out = torch.empty(m, 6 * mr, device=xyz.device, dtype=f32)
    w = (C.c_float * mr)(*window)
    call("occnerf_hann_pe", ptr(xyz, f32), m, C.cast(w, C.c_void_p), mr, ptr(out, f32), out.shape[1], stream())
    return out


# ----------------------------------------------------------------------------- non-rigid MLP on tensor cores
NR_PAIR = int(__import__("os").environ.get("OCCNERF_MLP_PAIR", "1")) != 0      # cta_group::2 pairs for the non-rigid chain as well


def nonrigid_pack(nr_w, nr_b, cond, n_pass, pair=None):
    """Packed UMMA operand images of the 7 non-rigid layers (cond (1,69) or None is folded into the first bias)."""
    pair = NR_PAIR if pair is None else pair
    dev = nr_w[0].device
    nbytes = _lib.load().occnerf_mlp_packed_bytes(n_pass, 2)
    packed = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    ws = [t.detach().contiguous().float() for t in nr_w]
    bs = [t.detach().contiguous().float() for t in nr_b]
    wp = (C.c_void_p * 7)(*[t.data_ptr() for t in ws])
    bp = (C.c_void_p * 7)(*[t.data_ptr() for t in bs])
    cond_c = cond.detach().reshape(-1).contiguous().float() if cond is not None else None
    call("occnerf_nonrigid_pack_weights", C.cast(wp, C.c_void_p), C.cast(bp, C.c_void_p), ptr(cond_c, f32), n_pass, int(pair), ptr(packed), stream())
    return packed


def nonrigid_forward_tc(xyz, window, packed, n_pass, out=None, pair=None):
    """xyz (m,3) -> xyz + non-rigid offsets, through the fused tcgen05 chain (csrc/mlp_tc.cu, chain 2); `pair` as given to nonrigid_pack."""
    pair = NR_PAIR if pair is None else pair
    m = xyz.shape[0]
    if len(window) != 6:
        raise RuntimeError("the fused non-rigid chain is built for 6 frequency bands (cfg.non_rigid_motion_mlp.multires)")
    if out is None:
        out = torch.empty(m, 3, device=xyz.device, dtype=f32)
    w = (C.c_float * 6)(*window)
    call("occnerf_nonrigid_forward_tc", ptr(xyz, f32), C.cast(w, C.c_void_p), m, ptr(packed), n_pass, int(pair), ptr(out, f32), stream())
    return out


# ----------------------------------------------------------------------------- K4 compositing
def composite_forward(raw, mask, z, rays, bg, want_weights=False, want_comp=False):
    N, S = z.shape
    dev = z.device
    rgb = torch.empty(N, 3, device=dev, dtype=f32)
    acc = torch.empty(N, device=dev, dtype=f32)
    depth = torch.empty(N, device=dev, dtype=f32)
    term = torch.empty(N, device=dev, dtype=i64)
    weights = torch.empty(N, S, device=dev, dtype=f32) if want_weights else None
    comp = torch.empty(N, S, device=dev, dtype=f32) if want_comp else None
    call("occnerf_composite_forward", ptr(raw, f32), ptr(mask, f32), ptr(z, f32), ptr(rays, f32), ptr(bg, f32), N, S,
         ptr(rgb), ptr(acc), ptr(depth), ptr(term), ptr(weights), ptr(comp), stream())
    return rgb, acc, depth, term, weights, comp


def composite_backward(raw, mask, z, rays, bg, g_rgb, g_acc, g_depth, g_comp=None):
    N, S = z.shape
    g_raw = torch.empty(N, S, 5, device=z.device, dtype=f32)
    g_mask = torch.empty(N, S, device=z.device, dtype=f32)
    # keep the (possibly freshly materialised) contiguous gradients alive until the launch has been issued
    g_rgb, g_acc, g_depth = g_rgb.contiguous(), g_acc.contiguous(), g_depth.contiguous()
    g_comp = g_comp.contiguous() if g_comp is not None else None
    call("occnerf_composite_backward", ptr(raw, f32), ptr(mask, f32), ptr(z, f32), ptr(rays, f32), ptr(bg, f32),
         ptr(g_rgb, f32), ptr(g_acc, f32), ptr(g_depth, f32), ptr(g_comp, f32), N, S, ptr(g_raw), ptr(g_mask), stream())
    return g_raw, g_mask


def visibility_hits(depth, term, x_skel, cloud4, k=10, thresh=0.5):
    N, S = x_skel.shape[0], x_skel.shape[1]
    V = cloud4.shape[0]
    hits = torch.empty(V, device=depth.device, dtype=f32)
    scratch = torch.empty(N * (k + 4) + 4, device=depth.device, dtype=i32)
    call("occnerf_visibility_hits", ptr(depth, f32), ptr(term, i64), ptr(x_skel, f32), N, S, float(thresh), ptr(cloud4, f32),
         V, k, ptr(hits), ptr(scratch), stream())
    return hits


# ----------------------------------------------------------------------------- rays in front of the path
def generate_rays(H, W, K, R, T, bbox_min, bbox_max, device="cuda", capacity=None, want_pixel_index=False, sync=True):
    """camera_util.py:133-160 + :163-212 + the dataset masking (freeview.py:208-219) on the device.

    K [3,3], R [3,3], T [3], bbox_min/max [3] are HOST arrays (numpy or CPU tensors) in the dtypes the reference's
    dataset holds them in; a float32 K makes `pixel_camera` float32 exactly as numpy would (the ZJU pickles), all-float32
    K, R, T (tpose.py:66-84) keep origin, directions and |d| in float32 as numpy does, everything else is float64.  Only these ~30 numbers cross to the device -- as kernel arguments.
    Returns (rays [n,8] float32 = (o, d, near, far) of the valid rays in pixel order -- the `ray_batch` layout of
    Network._render_rays --, ray_mask [H*W] bool, count, pixel_index [n] int32 or None).  With sync=False nothing is read
    back: rays has `capacity` rows (default H*W), of which the first count[0] (a device tensor) are valid.
    capacity=0 only fills ray_mask and the count; a capacity > 0 that is too small raises.
    """
    K = np.asarray(K)
    if K.dtype not in (np.float32, np.float64):
        raise RuntimeError(f"generate_rays: K must be float32 or float64 like the reference's cameras, got {K.dtype}")
    kinv = np.ascontiguousarray(np.linalg.inv(K).astype(np.float64))       # inverse in K's own dtype (camera_util.py:154)
    all_f32 = K.dtype == np.float32 and np.asarray(R).dtype == np.float32 and np.asarray(T).dtype == np.float32
    mode = 2 if all_f32 else (1 if K.dtype == np.float32 else 0)           # OCCNERF_RAYS_ALL_F32 / _K_F32 / _F64
    Rm = np.ascontiguousarray(np.asarray(R, dtype=np.float64).reshape(3, 3))
    Tv = np.ascontiguousarray(np.asarray(T, dtype=np.float64).reshape(3))
    lo = np.ascontiguousarray(np.asarray(bbox_min, dtype=np.float64).reshape(3))
    hi = np.ascontiguousarray(np.asarray(bbox_max, dtype=np.float64).reshape(3))
    H, W = int(H), int(W)
    cap = H * W if capacity is None else int(capacity)
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("occnerf_b200: generate_rays runs on a CUDA device (there is no CPU path)")
    with torch.cuda.device(dev):
        rays = torch.empty(cap, 8, device=dev, dtype=f32)
        mask = torch.empty(H * W, device=dev, dtype=u8)
        pix = torch.empty(cap, device=dev, dtype=i32) if want_pixel_index else None
        count = torch.empty(1, device=dev, dtype=i32)
        scratch = torch.empty(max(1, _lib.load().occnerf_rays_scratch_bytes(H, W) // 4), device=dev, dtype=i32)
        dp = lambda a: a.ctypes.data_as(C.c_void_p)
        call("occnerf_generate_rays", dp(kinv), mode, dp(Rm), dp(Tv), dp(lo), dp(hi), H, W, cap,
             ptr(rays, f32), ptr(mask, u8), ptr(pix, i32), ptr(count, i32), ptr(scratch, i32), stream())
    if not sync:
        return rays, mask.view(torch.bool), count, pix
    n = int(count.item())
    if cap == 0:                                   # mask / count query: nothing was asked to fit
        return rays, mask.view(torch.bool), n, pix
    if n > cap:
        raise RuntimeError(f"generate_rays: {n} valid rays do not fit the capacity of {cap}")
    return rays[:n], mask.view(torch.bool), n, (pix[:n] if pix is not None else None)


# ----------------------------------------------------------------------------- image assembly behind the path
# ----------------------------------------------------------------------------- training patches (csrc/patches.cu)
def draw_patch_randoms(n_patch, sample_subject_ratio, n_subject, n_bbox_minus_subject, rs=None):
    """The two draws per patch exactly as the reference takes them from numpy's global RandomState (train.py:198-201, 241-242):
    `rand(1)[0] < ratio`, then `choice(n_candidates, size=[1], replace=False)[0]`.  -> (use_subject bool [n], select_idx int [n])."""
    rs = np.random if rs is None else rs
    use, idx = [], []
    for _ in range(n_patch):
        u = bool(rs.rand(1)[0] < sample_subject_ratio)
        n = n_subject if u else n_bbox_minus_subject
        use.append(u)
        idx.append(int(rs.choice(n, size=[1], replace=False)[0]))
    return np.array(use), np.array(idx)


def sample_patches(ray_mask, subject_mask, bbox_mask, H, W, patch_size, use_subject, select_idx, rays=None):
    """train.py:167-273 on the device.  ray_mask / subject_mask / bbox_mask: uint8 or bool CUDA tensors with H*W elements;
    use_subject / select_idx: the caller's draws (array-likes, see draw_patch_randoms); rays [n_rays, 8] optional.
    -> dict(select_inds i32 [n*P*P] (first patch_div[-1] valid), patch_div i32 [n+1], patch_masks u8 [n,P,P], xy_min / xy_max i32 [n,2],
            rays [n*P*P, 8] | None, status i32 [1])."""
    dev = ray_mask.device
    as_u8 = lambda t: t.reshape(-1).to(torch.uint8).contiguous()
    rm, sm, bm = as_u8(ray_mask), as_u8(subject_mask), as_u8(bbox_mask)
    n, P = len(use_subject), int(patch_size)
    use_d = torch.as_tensor(np.asarray(use_subject, dtype=np.uint8), device=dev)
    idx_d = torch.as_tensor(np.asarray(select_idx, dtype=np.int32), device=dev)
    out = dict(select_inds=torch.empty(n * P * P, device=dev, dtype=i32), patch_div=torch.empty(n + 1, device=dev, dtype=i32),
               patch_masks=torch.empty(n, P, P, device=dev, dtype=u8), xy_min=torch.empty(n, 2, device=dev, dtype=i32),
               xy_max=torch.empty(n, 2, device=dev, dtype=i32), status=torch.zeros(1, device=dev, dtype=i32),
               rays=torch.empty(n * P * P, 8, device=dev, dtype=f32) if rays is not None else None)
    scratch = torch.empty(_lib.load().occnerf_patches_scratch_bytes(H, W, n, P), device=dev, dtype=u8)
    call("occnerf_sample_patches", ptr(rm, u8), ptr(sm, u8), ptr(bm, u8), H, W, P, n, ptr(use_d, u8), ptr(idx_d, i32),
         ptr(rays, f32) if rays is not None else None, ptr(out["rays"]) if rays is not None else None, ptr(out["select_inds"]),
         ptr(out["patch_div"]), ptr(out["patch_masks"]), ptr(out["xy_min"]), ptr(out["xy_max"]), ptr(out["status"]), ptr(scratch), stream())
    return out


class _PatchLoss(torch.autograd.Function):
    """trainer.py:31-41,24,135-189 without LPIPS: (rgb [n,3], comp_loss | None) -> (loss scalar, patch_imgs [N,P,P,3]); the gradients
    are produced by the same call (occnerf_patch_loss)."""

    @staticmethod
    def forward(ctx, rgb, comp, masks, div, bgcolor, targets, w_mse, w_comp):
        dev = rgb.device
        N, P = masks.shape[0], masks.shape[1]
        rgb_c = rgb.detach().contiguous().float()
        comp_c = comp.detach().contiguous().float() if comp is not None else None
        imgs = torch.empty(N, P, P, 3, device=dev, dtype=f32)
        out = torch.empty(4, device=dev, dtype=f32)
        g_rgb = torch.zeros_like(rgb_c)
        acc = torch.empty(2, device=dev, dtype=torch.float64)
        # (converted copies are held in locals until the call is enqueued: a temporary freed earlier could be handed out again and
        #  overwritten by the next conversion kernel on the same stream)
        masks_c, div_c = masks.to(u8).contiguous(), div.to(i32).contiguous()
        bg_c, tg_c = bgcolor.contiguous().float(), targets.contiguous().float()
        call("occnerf_patch_loss", ptr(rgb_c, f32), ptr(masks_c, u8), ptr(div_c, i32),
             ptr(bg_c, f32), ptr(tg_c, f32), ptr(comp_c, f32) if comp_c is not None else None,
             comp_c.numel() if comp_c is not None else 0, N, P, float(w_mse), float(w_comp), ptr(imgs), ptr(out), ptr(g_rgb), ptr(acc), stream())
        ctx.save_for_backward(g_rgb, out, masks_c)
        ctx.comp_shape = tuple(comp.shape) if comp is not None else None
        ctx.set_materialize_grads(False)
        parts = out[1:3].clone()
        ctx.mark_non_differentiable(parts)
        return out[0].clone(), imgs, parts

    @staticmethod
    def backward(ctx, g_loss, g_imgs, _g_parts):
        g_rgb, out, masks_c = ctx.saved_tensors
        g = g_rgb * g_loss if g_loss is not None else torch.zeros_like(g_rgb)
        if g_imgs is not None:              # a consumer of the unpacked images (the perceptual term): its gradient comes back through the scatter
            g = g + g_imgs[masks_c.bool()]  # (row-major order of the hit pixels = rank order; a host sync, outside the graph-captured step)
        g_comp = None
        if ctx.comp_shape is not None and g_loss is not None:
            g_comp = (g_loss * out[3]).expand(ctx.comp_shape).contiguous()
        return g, g_comp, None, None, None, None, None, None


def patch_loss(rgb, comp_loss, patch_masks, div_indices, bgcolor, targets, w_mse=0.2, w_comp=1.0):
    """-> (loss, patch_imgs, (w_mse * mse, w_comp * mean(comp_loss))); differentiable w.r.t. rgb and comp_loss."""
    return _PatchLoss.apply(rgb, comp_loss, patch_masks, div_indices, bgcolor, targets, w_mse, w_comp)


def unpack_image(rgb, alpha, pixel_index, H, W, bgcolor, out=None, fill=True):
    """run.py:39-66 (unpack_to_image / unpack_alpha_map) + image_util.py:19-20 (to_8b_image) on the device.

    rgb [n,3], alpha [n] or None, pixel_index [n] int32 (from generate_rays); bgcolor: three HOST numbers in [0,1]
    (run.py:118 passes cfg.bgcolor / 255).  Returns (rgb8 [H,W,3] uint8, alpha8 [H,W] uint8, bad [1] int32 = number of
    pixel indices outside the frame, a device tensor so that nothing is read back here).  `out=(rgb8, alpha8)` with
    fill=False scatters a further shard of rays into an already painted frame."""
    n = int(rgb.shape[0])
    dev = rgb.device
    if out is None:
        if not fill:
            raise RuntimeError("unpack_image: fill=False needs the frame to scatter into (out=...)")
        rgb8 = torch.empty(H, W, 3, device=dev, dtype=u8)
        alpha8 = torch.empty(H, W, device=dev, dtype=u8)
    else:
        rgb8, alpha8 = out
        if tuple(rgb8.shape) != (H, W, 3) or (alpha8 is not None and tuple(alpha8.shape) != (H, W)):
            raise RuntimeError(f"unpack_image: out has shapes {tuple(rgb8.shape)} / {None if alpha8 is None else tuple(alpha8.shape)}, "
                               f"expected ({H}, {W}, 3) / ({H}, {W})")
    bg = np.ascontiguousarray(np.asarray(bgcolor, dtype=np.float64).reshape(3).astype(np.float32))   # np.full(..., dtype='float32')
    bad = torch.zeros(1, device=dev, dtype=i32)
    call("occnerf_unpack_image", ptr(rgb, f32) if n else None, ptr(alpha, f32) if (n and alpha is not None) else None,
         ptr(pixel_index, i32) if n else None, n, int(H), int(W), bg.ctypes.data_as(C.c_void_p), 1 if fill else 0,
         ptr(rgb8, u8), ptr(alpha8, u8), ptr(bad, i32), stream())
    return rgb8, alpha8, bad
