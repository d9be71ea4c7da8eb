"""Seeded synthetic inputs shaped like the reference's ZJU-Mocap batches.

Nothing here is on the timed path: it builds the tensors the hot path consumes
(`rays, near, far, dst_Rs, dst_Ts, cnl_gtfms, motion_weights_priors, dst_posevec, cnl_bbox_*`),
the per-subject state (`point_base, point_norms, fps_index`) and random-init weights.  The real
SMPL template is not redistributable (third_parties/smpl/models/PUT_SMPL_MODEL_HERE in the
reference), so vertices are capsule surfaces around a 24-joint T-pose skeleton.

Follows, without importing them:
  core/utils/body_util.py:222-371   (FK, canonical transforms, Gaussian bone volumes)
  core/utils/camera_util.py:133-212 (pinhole rays, ray/box near-far)
  core/utils/network_util.py:138-200 (motion basis), :207-334 (xavier `initseq`)
  core/data/occnerf/tpose.py:22-84  (synthetic camera), train.py:225-273 (patch rays)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

TOTAL_BONES = 24
PARENT = {1: 0, 2: 0, 3: 0, 4: 1, 5: 2, 6: 3, 7: 4, 8: 5, 9: 6, 10: 7, 11: 8, 12: 9, 13: 9,
          14: 9, 15: 12, 16: 13, 17: 14, 18: 16, 19: 17, 20: 18, 21: 19, 22: 20, 23: 21}
TORSO = (0, 3, 6, 9, 13, 14)
HEAD = 15

# SMPL-like neutral T-pose joints (metres); x right-left, y up, z front.
TPOSE_JOINTS = np.array([
    [0.000, -0.240, 0.030], [0.070, -0.330, 0.020], [-0.070, -0.330, 0.020], [0.000, -0.130, 0.000],
    [0.105, -0.715, 0.015], [-0.105, -0.715, 0.015], [0.000, 0.010, 0.010], [0.090, -1.115, -0.030],
    [-0.090, -1.115, -0.030], [0.000, 0.065, 0.030], [0.115, -1.175, 0.095], [-0.115, -1.175, 0.095],
    [0.000, 0.275, -0.010], [0.080, 0.180, -0.010], [-0.080, 0.180, -0.010], [0.000, 0.360, 0.040],
    [0.170, 0.225, -0.020], [-0.170, 0.225, -0.020], [0.430, 0.210, -0.040], [-0.430, 0.210, -0.040],
    [0.680, 0.215, -0.045], [-0.680, 0.215, -0.045], [0.765, 0.205, -0.055], [-0.765, 0.205, -0.055],
], dtype=np.float32)


# ----------------------------------------------------------------------------- skeleton / pose
def rodrigues(rvec: np.ndarray) -> np.ndarray:
    """Axis-angle -> 3x3, with the reference's (norm + 1e-5) normalisation (body_util.py:201-219)."""
    rvec = np.asarray(rvec, dtype=np.float64).reshape(3)
    theta = np.linalg.norm(rvec)
    r = rvec / (theta + 1e-5)
    K = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
    return math.cos(theta) * np.eye(3) + math.sin(theta) * K + (1 - math.cos(theta)) * np.outer(r, r)


def pose_to_local_RTs(pose72: np.ndarray, joints: np.ndarray):
    """body_util.py:222-248: per-joint local rotation and parent-relative offset."""
    ang = pose72.reshape(-1, 3)
    Rs = np.stack([rodrigues(a) for a in ang]).astype(np.float32)
    Ts = joints.copy().astype(np.float32)
    for i in range(1, TOTAL_BONES):
        Ts[i] = joints[i] - joints[PARENT[i]]
    return Rs, Ts


def canonical_global_tfms(joints: np.ndarray) -> np.ndarray:
    """body_util.py:251-271: rest-pose global 4x4 of every joint (identity rotations)."""
    G = np.zeros((TOTAL_BONES, 4, 4), dtype=np.float32)
    for i in range(TOTAL_BONES):
        G[i] = np.eye(4, dtype=np.float32)
        G[i, :3, 3] = joints[i]
    return G


def motion_basis(dst_Rs: np.ndarray, dst_Ts: np.ndarray, cnl_gtfms: np.ndarray):
    """network_util.py:166-200: f_i = G_cnl,i . inverse(G_dst,i) -> (R_i, T_i), float32 like the reference."""
    local = torch.zeros(TOTAL_BONES, 4, 4)
    local[:, :3, :3] = torch.from_numpy(dst_Rs)
    local[:, :3, 3] = torch.from_numpy(dst_Ts)
    local[:, 3, 3] = 1.0
    glob = torch.zeros_like(local)
    glob[0] = local[0]
    for i in range(1, TOTAL_BONES):
        glob[i] = glob[PARENT[i]] @ local[i]
    f = torch.from_numpy(cnl_gtfms) @ torch.inverse(glob)
    return f[:, :3, :3].contiguous(), f[:, :3, 3].contiguous(), glob


def skeleton_bbox(joints: np.ndarray, offset: float = 0.3):
    return (joints.min(0) - offset).astype(np.float32), (joints.max(0) + offset).astype(np.float32)


# ----------------------------------------------------------------------------- motion-weight prior
def _rot_between(v1: np.ndarray, v2: np.ndarray) -> np.ndarray:
    v1 = v1 / max(np.linalg.norm(v1), 1e-5)
    v2 = v2 / max(np.linalg.norm(v2), 1e-5)
    n = np.cross(v1, v2)
    c = float(v1 @ v2)
    K = np.array([[0, -n[2], n[1]], [n[2], 0, -n[0]], [-n[1], n[0], 0]])
    return (np.eye(3) + K + K @ K * (1.0 / (1.0 + c))).astype(np.float32)


def gaussian_bone_volumes(joints: np.ndarray, bmin, bmax, grid: int = 32) -> np.ndarray:
    """body_util.py:274-371: one anisotropic Gaussian per bone (+ background), normalised over channels."""
    zs, ys, xs = np.meshgrid(np.linspace(bmin[2], bmax[2], grid), np.linspace(bmin[1], bmax[1], grid),
                             np.linspace(bmin[0], bmax[0], grid), indexing="ij")

    def blob(center, S, R):
        sigma = R @ S @ S @ R.T
        g = np.stack([xs - center[0], ys - center[1], zs - center[2]], -1)
        return np.exp(-np.einsum("abci,ij,abcj->abc", g, sigma, g))

    def scale(stds):
        return np.diag(1.0 / np.asarray(stds, dtype=np.float32))

    vols = []
    for j in range(TOTAL_BONES):
        v = np.zeros((grid,) * 3, dtype=np.float32)
        children = [c for c, p in PARENT.items() if p == j]
        for c in children:
            S = scale(np.array([0.03, 0.06, 0.03]) * 2.0)
            if j in TORSO:
                S[0, 0] /= 1.5
                S[2, 2] /= 1.5
            a, b = joints[PARENT[c]], joints[c]
            v = v + blob((a + b) / 2.0, S, _rot_between(np.array([0.0, 1.0, 0.0]), b - a))
        if not children:
            stds = np.array([0.06] * 3) if j == HEAD else np.array([0.02] * 3)
            v = blob(joints[j], scale(stds * 2.0), np.eye(3, dtype=np.float32))
        vols.append(v)
    vols = np.stack(vols, 0)
    bg = 1.0 - vols.sum(0, keepdims=True).clip(0.0, 1.0)
    vols = np.concatenate([vols, bg], 0)
    return (vols / vols.sum(0, keepdims=True).clip(min=0.001)).astype(np.float32)


# ----------------------------------------------------------------------------- vertices
def capsule_vertices(joints: np.ndarray, V: int = 6890, seed: int = 0):
    """Stand-in for SMPL vertices + trimesh vertex normals (network.py:92-98): points on capsule
    surfaces around every bone with analytic outward normals."""
    rng = np.random.default_rng(seed)
    radius = {0: 0.12, 3: 0.13, 6: 0.14, 9: 0.14, 12: 0.06, 13: 0.07, 14: 0.07, 15: 0.10,
              1: 0.075, 2: 0.075, 4: 0.055, 5: 0.055, 7: 0.045, 8: 0.045, 10: 0.035, 11: 0.035,
              16: 0.05, 17: 0.05, 18: 0.04, 19: 0.04, 20: 0.03, 21: 0.03, 22: 0.025, 23: 0.025}
    bones = [(PARENT[c], c) for c in sorted(PARENT)]
    lens = np.array([np.linalg.norm(joints[c] - joints[p]) + 2 * radius[c] for p, c in bones])
    area = lens * np.array([radius[c] for _, c in bones])
    counts = np.floor(area / area.sum() * V).astype(int)
    counts[0] += V - counts.sum()
    pts, nrm = [], []
    for (p, c), n in zip(bones, counts):
        a, b, r = joints[p].astype(np.float64), joints[c].astype(np.float64), radius[c]
        axis = b - a
        L = np.linalg.norm(axis)
        axis = axis / L
        t = rng.uniform(-r, L + r, n)                     # position along the capsule axis incl. caps
        u = rng.normal(size=(n, 3))
        u -= np.outer(u @ axis, axis)
        u /= np.linalg.norm(u, axis=1, keepdims=True)
        tc = np.clip(t, 0.0, L)
        over = t - tc                                     # how far into a hemispherical cap
        rad = np.sqrt(np.maximum(r * r - over * over, 0.0))
        centre = a + np.outer(tc, axis)
        x = centre + np.outer(over, axis) + u * rad[:, None]
        nn = x - centre
        nn /= np.maximum(np.linalg.norm(nn, axis=1, keepdims=True), 1e-9)
        pts.append(x)
        nrm.append(nn)
    return np.concatenate(pts).astype(np.float32), np.concatenate(nrm).astype(np.float32)


def farthest_point_sampling(points: np.ndarray, n: int) -> np.ndarray:
    """Deterministic greedy FPS from index 0 (stand-in for torch_cluster.fps, network.py:113-118)."""
    P = points.astype(np.float64)
    sel = np.zeros(n, dtype=np.int64)
    d = np.full(P.shape[0], np.inf)
    cur = 0
    for i in range(n):
        sel[i] = cur
        d = np.minimum(d, ((P - P[cur]) ** 2).sum(1))
        cur = int(np.argmax(d))
    return sel


# ----------------------------------------------------------------------------- camera / rays
def lookat_camera(img: int, campos=(0.0, -0.25, 6.0), focal_512: float = 1250.0, yaw: float = 0.0):
    """tpose.py:66-84 + camera_util.py:40-83 (`inv_camera=True`); `yaw` rotates the camera position about
    the vertical axis through the look-at point (freeview.py:133-142)."""
    campos = np.asarray(campos, dtype=np.float32).copy()
    look = np.array([0.0, campos[1], 0.0], dtype=np.float32)
    c, s = math.cos(yaw), math.sin(yaw)
    campos = np.array([c * campos[0] + s * campos[2], campos[1], -s * campos[0] + c * campos[2]], dtype=np.float32)
    up = np.array([0.0, -1.0, 0.0], dtype=np.float32)
    fwd = look - campos
    fwd /= np.linalg.norm(fwd)
    right = np.cross(up, fwd)
    right /= np.linalg.norm(right)
    up = np.cross(fwd, right)
    up /= np.linalg.norm(up)
    R = np.stack([right, up, fwd]).astype(np.float32)
    T = -R @ campos
    K = np.eye(3, dtype=np.float32)
    K[0, 0] = K[1, 1] = focal_512 * img / 512.0
    K[:2, 2] = img / 2.0
    return K, R, T


def pixel_rays(H: int, W: int, K, R, T):
    """camera_util.py:133-160: un-normalised world-space ray directions through pixel centres (integer grid)."""
    o = -(R.T @ T)
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    cam = np.stack([i, j, np.ones_like(i)], 2) @ np.linalg.inv(K).T
    world = (cam - T) @ R
    d = (world - o).astype(np.float32)
    return np.broadcast_to(o.astype(np.float32), d.shape).copy(), d


def ray_box_near_far(bmin, bmax, o: np.ndarray, d: np.ndarray):
    """camera_util.py:163-212: slab test returning near/far in units of |d| and the hit mask."""
    lo = np.asarray(bmin, np.float64) - 0.01
    hi = np.asarray(bmax, np.float64) + 0.01
    d = d.astype(np.float64).copy()
    d[np.abs(d) < 1e-5] = 1e-5
    o = o.astype(np.float64)
    t0 = (lo - o) / d
    t1 = (hi - o) / d
    tn = np.minimum(t0, t1).max(1)
    tf = np.maximum(t0, t1).min(1)
    hit = tf > tn
    return tn[hit].astype(np.float32), tf[hit].astype(np.float32), hit


# ----------------------------------------------------------------------------- weights
def _xavier_uniform(shape, fan_sum, gain, gen):
    bound = gain * math.sqrt(2.0 / fan_sum) * math.sqrt(3.0)
    return (torch.rand(shape, generator=gen) * 2 - 1) * bound


def init_mlp(dims, gen, relu_last=False):
    """`initseq` (network_util.py:316-334): xavier-uniform, gain sqrt(2) when followed by ReLU, zero bias."""
    ws, bs = [], []
    for li, (i, o) in enumerate(zip(dims[:-1], dims[1:])):
        followed_by_relu = relu_last or li < len(dims) - 2
        ws.append(_xavier_uniform((o, i), i + o, math.sqrt(2.0) if followed_by_relu else 1.0, gen))
        bs.append(torch.zeros(o))
    return ws, bs


def hashgrid_offsets(input_dim=4, num_levels=16, base_resolution=16, log2_hashmap_size=19,
                     desired_resolution=None, per_level_scale=2.0):
    """grid.py:102-135: per-level table sizes (dense until (res+1)^D exceeds 2^19, rounded up to 8)."""
    if desired_resolution is not None:
        per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    offs, off = [], 0
    for i in range(num_levels):
        res = int(np.ceil(base_resolution * per_level_scale ** i))
        n = min(2 ** log2_hashmap_size, (res + 1) ** input_dim)
        n = int(np.ceil(n / 8) * 8)
        offs.append(off)
        off += n
    offs.append(off)
    return np.array(offs, dtype=np.int32), float(per_level_scale)


@dataclass
class Subject:
    """Per-subject state the reference builds in `generate_neural_points` (network.py:90-146)."""
    joints: np.ndarray
    bbox_min: np.ndarray
    bbox_max: np.ndarray
    bound: float
    point_base: torch.Tensor          # (V,3)
    point_norms: torch.Tensor         # (V,3)
    point_dist: torch.Tensor          # (V,1) learnable, U(+-1e-4)
    point_counter: torch.Tensor       # (V,)
    fps_index: list                   # 3 x LongTensor (V/4, V/16, V/64)
    priors: torch.Tensor              # (25,32,32,32)


@dataclass
class NetWeights:
    """Random-init parameters with the reference's names and shapes."""
    embeddings: torch.Tensor          # (7755336, 2)
    offsets: torch.Tensor             # (17,) int32
    per_level_scale: float
    pts_w: list = field(default_factory=list)   # pts_linears.{0,2,4,6}.weight
    pts_b: list = field(default_factory=list)
    geo_w: torch.Tensor = None        # geo_linear.0  (65,256)
    geo_b: torch.Tensor = None
    rgb_w: list = field(default_factory=list)   # rgb_linears.{0,2,4,6}
    rgb_b: list = field(default_factory=list)
    out_w: torch.Tensor = None        # output_linear.0 (3,256)
    out_b: torch.Tensor = None
    nr_w: list = field(default_factory=list)    # non_rigid_mlp block_mlps.{0,..,12}
    nr_b: list = field(default_factory=list)


def make_subject(seed: int = 0, V: int = 6890, bbox_offset: float = 0.3) -> Subject:
    joints = TPOSE_JOINTS.copy()
    bmin, bmax = skeleton_bbox(joints, bbox_offset)
    bound = float(np.max(np.abs(np.concatenate([bmin, bmax]))))
    verts, norms = capsule_vertices(joints, V, seed)
    gen = torch.Generator().manual_seed(seed + 17)
    fps = []
    ratio = 1.0
    for _ in range(3):
        ratio /= 4
        fps.append(torch.from_numpy(farthest_point_sampling(verts, int(math.ceil(V * ratio)))))
    return Subject(
        joints=joints, bbox_min=bmin, bbox_max=bmax, bound=bound,
        point_base=torch.from_numpy(verts), point_norms=torch.from_numpy(norms),
        point_dist=(torch.rand(V, 1, generator=gen) * 2 - 1) * 1e-4,
        point_counter=torch.ones(V), fps_index=fps,
        priors=torch.from_numpy(gaussian_bone_volumes(joints, bmin, bmax, 32)))


def make_weights(bound: float, seed: int = 0, table_scale: float = 1e-4, nonzero_bias: bool = False) -> NetWeights:
    """`table_scale`/`nonzero_bias` let tests leave the random-init regime (hash features 1e-4, zero
    biases) where many code paths are numerically invisible."""
    gen = torch.Generator().manual_seed(seed + 101)
    offs, pls = hashgrid_offsets(desired_resolution=2048 * bound)
    emb = (torch.rand(int(offs[-1]), 2, generator=gen) * 2 - 1) * table_scale
    w = NetWeights(embeddings=emb, offsets=torch.from_numpy(offs), per_level_scale=pls)
    w.pts_w, w.pts_b = init_mlp([68, 256, 256, 256, 256], gen, relu_last=True)
    (w.geo_w,), (w.geo_b,) = init_mlp([256, 65], gen)
    w.rgb_w, w.rgb_b = init_mlp([131, 256, 256, 256, 256], gen, relu_last=True)
    (w.out_w,), (w.out_b,) = init_mlp([256, 3], gen)
    # non-rigid MLP (mlp_offset.py:16-42): 105->128 x4, (128+36)->128, 128->128, 128->3 (U(+-1e-5), zero bias)
    dims_in = [105, 128, 128, 128, 164, 128]
    for i, di in enumerate(dims_in):
        w.nr_w.append(_xavier_uniform((128, di), di + 128, math.sqrt(2.0), gen))
        w.nr_b.append(torch.zeros(128))
    w.nr_w.append((torch.rand(3, 128, generator=gen) * 2 - 1) * 1e-5)
    w.nr_b.append(torch.zeros(3))
    if nonzero_bias:
        for lst in (w.pts_b, w.rgb_b, w.nr_b):
            for b in lst:
                b.copy_((torch.rand(b.shape, generator=gen) * 2 - 1) * 0.05)
        w.geo_b.copy_((torch.rand(65, generator=gen) * 2 - 1) * 0.05)
        w.out_b.copy_((torch.rand(3, generator=gen) * 2 - 1) * 0.05)
    return w


def make_motion_weights_vol(priors: torch.Tensor, seed: int = 0, logit_std: float = 0.5) -> torch.Tensor:
    """Stand-in for `MotionWeightVolumeDecoder` at random init (deconv_vol_decoder.py:25-33):
    softmax(smooth random logits + log prior) over the 25 channels.  The real decoder (63.6 M
    parameters) is outside the hot path; only its output distribution
    matters here: strictly positive where the prior is, summing to one."""
    gen = torch.Generator().manual_seed(seed + 303)
    coarse = torch.randn(1, 25, 8, 8, 8, generator=gen) * logit_std
    logits = torch.nn.functional.interpolate(coarse, size=(32, 32, 32), mode="trilinear", align_corners=True)[0]
    return torch.softmax(logits + torch.log(priors), dim=0).contiguous()


@dataclass
class Frame:
    """One `Network.forward` call worth of inputs (network.py:542-549)."""
    rays_o: torch.Tensor
    rays_d: torch.Tensor
    near: torch.Tensor
    far: torch.Tensor
    dst_Rs: torch.Tensor
    dst_Ts: torch.Tensor
    cnl_gtfms: torch.Tensor
    motion_scale_Rs: torch.Tensor     # (24,3,3)  output of the motion-basis prologue
    motion_Ts: torch.Tensor           # (24,3)
    dst_posevec: torch.Tensor         # (69,)
    cnl_bbox_min_xyz: torch.Tensor
    cnl_bbox_scale_xyz: torch.Tensor
    bgcolor: torch.Tensor
    ray_mask: np.ndarray = None
    img_hw: tuple = None


def make_frame(subject: Subject, mode: str = "patch", img: int = 512, n_patches: int = 6, patch: int = 32,
               seed: int = 0, yaw: float = 0.0, pose_std: float = 0.2, max_rays: int | None = None,
               subject_ratio: float = 0.8, bbox_offset: float = 0.3, occlusion_band: tuple | None = None) -> Frame:
    """bbox_offset: margin of the posed box the rays are clipped to (cfg.bbox_offset: 0.3 for ZJU-Mocap, 2.0 in the OcMotion yamls);
    occlusion_band = (mid, width) in pixels: the column band the dataset blanks in the alpha mask (core/data/occnerf/train.py:286-287),
    patches are not placed on it."""
    rng = np.random.default_rng(seed + 7)
    pose = np.zeros(72, dtype=np.float32)
    pose[3:] = rng.normal(0.0, pose_std, 69).astype(np.float32)
    Rs, Ts = pose_to_local_RTs(pose, subject.joints)
    G = canonical_global_tfms(subject.joints)
    mRs, mTs, glob = motion_basis(Rs, Ts, G)
    posed_joints = glob[:, :3, 3].numpy()
    dmin, dmax = skeleton_bbox(posed_joints, bbox_offset)
    K, R, T = lookat_camera(img, yaw=yaw)
    o, d = pixel_rays(img, img, K, R, T)
    o, d = o.reshape(-1, 3), d.reshape(-1, 3)
    near, far, hit = ray_box_near_far(dmin, dmax, o, d)
    hit2d = hit.reshape(img, img)
    if mode == "patch":
        ys, xs = np.nonzero(hit2d)
        sel = np.zeros_like(hit2d)
        order = np.full(hit2d.shape, -1, dtype=np.int64)
        count, tries = 0, 0
        bones = [(PARENT[c], c) for c in sorted(PARENT)]
        while count < n_patches and tries < 10000:
            tries += 1
            if rng.uniform() < subject_ratio:
                # patch centred on the projected skeleton (train.py:225-273 `sample_subject_ratio`)
                a, b = bones[rng.integers(len(bones))]
                X = posed_joints[a] + rng.uniform() * (posed_joints[b] - posed_joints[a])
                uvw = K @ (R @ X + T)
                cx, cy = int(round(uvw[0] / uvw[2])), int(round(uvw[1] / uvw[2]))
            else:
                k = rng.integers(len(ys))
                cy, cx = ys[k], xs[k]
            y0, x0 = cy - patch // 2, cx - patch // 2
            if y0 < 0 or x0 < 0 or y0 + patch > img or x0 + patch > img:
                continue
            if not hit2d[y0:y0 + patch, x0:x0 + patch].all() or sel[y0:y0 + patch, x0:x0 + patch].any():
                continue
            if occlusion_band is not None and abs(cx - occlusion_band[0]) < occlusion_band[1] // 2:
                continue
            sel[y0:y0 + patch, x0:x0 + patch] = True
            order[y0:y0 + patch, x0:x0 + patch] = count * patch * patch + np.arange(patch * patch).reshape(patch, patch)
            count += 1
        assert count == n_patches, "could not place the requested patches inside the box silhouette"
        full_idx = np.nonzero(sel.reshape(-1))[0]
        full_idx = full_idx[np.argsort(order.reshape(-1)[full_idx])]       # patch-major ray order
        hit_rank = np.cumsum(hit) - 1
        keep = hit_rank[full_idx]
        o, d, near, far = o[full_idx], d[full_idx], near[keep], far[keep]
    else:
        o, d = o[hit], d[hit]
    if max_rays is not None and o.shape[0] > max_rays:
        pick = np.sort(rng.choice(o.shape[0], max_rays, replace=False))
        o, d, near, far = o[pick], d[pick], near[pick], far[pick]
    cmin, cmax = subject.bbox_min, subject.bbox_max
    return Frame(
        rays_o=torch.from_numpy(o.copy()), rays_d=torch.from_numpy(d.copy()),
        near=torch.from_numpy(near.copy())[:, None], far=torch.from_numpy(far.copy())[:, None],
        dst_Rs=torch.from_numpy(Rs), dst_Ts=torch.from_numpy(Ts), cnl_gtfms=torch.from_numpy(G),
        motion_scale_Rs=mRs, motion_Ts=mTs, dst_posevec=torch.from_numpy(pose[3:] + 1e-2),
        cnl_bbox_min_xyz=torch.from_numpy(cmin.copy()), cnl_bbox_scale_xyz=torch.from_numpy((2.0 / (cmax - cmin)).astype(np.float32)),
        bgcolor=torch.zeros(3), ray_mask=hit, img_hw=(img, img))


def network_from_synthetic(subject: Subject, weights: NetWeights, cfg=None, device="cuda"):
    """An occnerf_b200.network.Network carrying the synthetic per-subject state and the given parameters
    (the counterpart of oracle/ref_shim.build_reference_network for the CUDA path)."""
    from occnerf_b200.network import Network
    net = Network(cfg)
    net.generate_neural_points(subject.point_base, subject.point_norms, subject.fps_index, subject.bbox_min,
                               subject.bbox_max, point_dist=subject.point_dist, point_counter=subject.point_counter)
    assert abs(net.bound - subject.bound) < 1e-6
    net.bound = subject.bound
    m = net.cnl_mlp.module
    with torch.no_grad():
        assert torch.equal(m.encoder.offsets, weights.offsets), "hash-grid level table differs from the reference's"
        m.encoder.embeddings.copy_(weights.embeddings)
        for i, li in enumerate((0, 2, 4, 6)):
            m.pts_linears[li].weight.copy_(weights.pts_w[i]); m.pts_linears[li].bias.copy_(weights.pts_b[i])
            m.rgb_linears[li].weight.copy_(weights.rgb_w[i]); m.rgb_linears[li].bias.copy_(weights.rgb_b[i])
        m.geo_linear[0].weight.copy_(weights.geo_w); m.geo_linear[0].bias.copy_(weights.geo_b)
        m.output_linear[0].weight.copy_(weights.out_w); m.output_linear[0].bias.copy_(weights.out_b)
        nr = net.non_rigid_mlp.module
        for i, li in enumerate(range(0, 14, 2)):
            nr.block_mlps[li].weight.copy_(weights.nr_w[i]); nr.block_mlps[li].bias.copy_(weights.nr_b[i])
    return net.to(device)


def frame_to(frame: Frame, device):
    """Moves the tensor fields of a Frame to `device` (numpy fields stay on the host)."""
    import dataclasses
    kw = {}
    for f in dataclasses.fields(frame):
        v = getattr(frame, f.name)
        kw[f.name] = v.to(device) if torch.is_tensor(v) else v
    return Frame(**kw)
