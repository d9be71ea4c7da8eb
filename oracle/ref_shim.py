"""ORACLE (test infrastructure): run the UNMODIFIED reference hot path on CPU, in this container only.

`/root/reference` cannot be imported as is under Python 3.12 / CPU-only torch:
`imp` is gone, `configs` parses argv and counts GPUs at import, and pykeops / torch_cluster / trimesh /
pytorch3d / termcolor and the two CUDA extensions are absent.  This module installs just enough fakes
for `core.nets.occnerf.network.Network` to import and for `_render_rays` / `_query_mlp` /
`_raw2outputs` / `_sample_motion_fields` to run unmodified, with two CPU stand-ins where the reference
has no CPU code at all:
  * `_gridencoder`  -> oracle/hashgrid_oracle.c (restating gridencoder.cu)
  * `pykeops.torch.LazyTensor` -> exact brute-force K-min (restating the call pattern of knn.py:33-85)

It is used by oracle/make_golden.py to write tests/golden/*.npz; nothing that runs on the GPU box
imports it (the reference does not exist there).
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("OCCNERF_REFERENCE_ROOT", "/root/reference")
REF_CFG = "configs/occnerf/zju_mocap/387/occnerf.yaml"

_loaded = {}


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "core", "nets", "occnerf"))


# --------------------------------------------------------------------------- fake third-party modules
def _sqdist_f32(q: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    """(dx*dx + dy*dy) + dz*dz in float32 with no contraction -- the distance definition every KNN in
    this repo (oracle and CUDA) shares.  KeOps itself ranks sqrt of this; sqrt is monotone."""
    dx = q[:, None, 0] - s[None, :, 0]
    dy = q[:, None, 1] - s[None, :, 1]
    dz = q[:, None, 2] - s[None, :, 2]
    return (dx * dx + dy * dy) + dz * dz


def exact_kmin(q: torch.Tensor, s: torch.Tensor, k: int, chunk: int = 4096) -> torch.Tensor:
    """k smallest distances, ascending, ties -> lowest support index.  Returns int64 (n,k)."""
    out = torch.empty(q.shape[0], k, dtype=torch.int64)
    for i in range(0, q.shape[0], chunk):
        d = _sqdist_f32(q[i:i + chunk].float(), s.float())
        out[i:i + chunk] = torch.sort(d, dim=1, stable=True)[1][:, :k]
    return out


class _FakeLazyTensor:
    def __init__(self, t):
        self.t = t.detach()
        self.ranges = None

    def __sub__(self, other):
        r = _FakeLazyTensor(self.t)
        r.other = other.t
        return r

    def norm2(self):
        return self

    def Kmin_argKmin(self, k, dim):
        q = self.t.reshape(-1, self.t.shape[-1])
        s = self.other.reshape(-1, self.other.shape[-1])
        if self.ranges is None:
            idx = exact_kmin(q, s, k)
            dist = (q[:, None, :].float() - s[idx].float()).norm(dim=-1)
            return dist, idx
        else:
            ranges_x, _sx, ranges_y = self.ranges[0], self.ranges[1], self.ranges[2]
            idx = torch.empty(q.shape[0], k, dtype=torch.int64)
            for (x0, x1), (y0, y1) in zip(ranges_x.tolist(), ranges_y.tolist()):
                idx[x0:x1] = exact_kmin(q[x0:x1], s[y0:y1], k) + y0
        return None, idx


def _install_fakes():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    def load_source(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    mod("imp", load_source=load_source)
    mod("trimesh", Trimesh=object)
    mod("torch_cluster", fps=None)
    mod("pytorch3d")
    mod("pytorch3d.ops")
    mod("pytorch3d.ops.points_normals", estimate_pointcloud_normals=None)
    mod("pykeops")
    mod("pykeops.torch", LazyTensor=_FakeLazyTensor)
    mod("termcolor", colored=lambda s, *a, **k: s)
    mod("_shencoder")
    from oracle import hashgrid_c
    mod("_gridencoder", grid_encode_forward=hashgrid_c.RefBackendModule.grid_encode_forward,
        grid_encode_backward=hashgrid_c.RefBackendModule.grid_encode_backward,
        grad_total_variation=hashgrid_c.RefBackendModule.grad_total_variation)


def load_reference():
    """Import the reference `network` module; returns (module, cfg).  Process-global and idempotent."""
    if "net_mod" in _loaded:
        return _loaded["net_mod"], _loaded["cfg"]
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    _install_fakes()
    old_argv, old_cwd, old_dc = sys.argv, os.getcwd(), torch.cuda.device_count
    sys.path.insert(0, REF_ROOT)
    os.chdir(REF_ROOT)
    sys.argv = ["ref_shim", "--cfg", REF_CFG]
    torch.cuda.device_count = lambda: 1
    try:
        with contextlib.redirect_stdout(open(os.devnull, "w")):
            from configs import cfg
            import core.nets.occnerf.network as net_mod
    finally:
        torch.cuda.device_count = old_dc
        sys.argv = old_argv
    # stay in REF_ROOT: component_factory resolves module paths relative to cwd at Network() time
    _loaded["cwd"] = old_cwd
    cfg.primary_gpus = ["cpu"]
    cfg.secondary_gpus = ["cpu"]
    _loaded.update(net_mod=net_mod, cfg=cfg)
    return net_mod, cfg


def build_reference_network(subject, weights):
    """A reference `Network` whose per-subject state and parameters are the synthetic ones.

    `generate_neural_points` (network.py:90-146) needs the SMPL pkl and trimesh, so the attribute
    set-up it performs is reproduced here with synthetic vertices; everything downstream is the
    reference's own code.
    """
    net_mod, cfg = load_reference()
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        with contextlib.redirect_stdout(open(os.devnull, "w")):
            net = net_mod.Network()
            V = subject.point_base.shape[0]
            net.bound = subject.bound
            net.detailed_bound = torch.tensor([list(subject.bbox_min), list(subject.bbox_max)])
            net.point_base = nn.Parameter(subject.point_base.clone().float(), requires_grad=False)
            net.point_dist = nn.Parameter(subject.point_dist.clone().float(), requires_grad=True)
            net.fps_index = [f.clone() for f in subject.fps_index]
            net.point_counter = nn.Parameter(subject.point_counter.clone().float(), requires_grad=False)
            net.point_norms = subject.point_norms.clone()
            ranges = torch.cumsum(torch.tensor([0, V] + [f.shape[0] for f in net.fps_index]), 0)
            net.ranges_y = torch.stack((ranges[:-1], ranges[1:])).t().int().contiguous()
            net.slices_x = (torch.arange(0, 4) + 1).int().view(4,)
            net.slices_y = (torch.arange(0, 4) + 1).int().view(4,)
            net.offset = ranges[:-1].view(4, 1)
            from core.nets.occnerf.component_factory import load_canonical_mlp
            mlp = load_canonical_mlp(cfg.canonical_mlp.module)(
                input_ch=net.cnl_pos_embed_size, mlp_depth=cfg.canonical_mlp.mlp_depth,
                mlp_width=cfg.canonical_mlp.mlp_width, skips=[], bound=net.bound, detailed_bound=net.detailed_bound)
            net.cnl_mlp = nn.DataParallel(mlp, device_ids=cfg.secondary_gpus, output_device=cfg.primary_gpus[0])
    finally:
        os.chdir(cwd)
    m = net.cnl_mlp.module
    with torch.no_grad():
        assert m.encoder.embeddings.shape == weights.embeddings.shape, (m.encoder.embeddings.shape, weights.embeddings.shape)
        assert torch.equal(m.encoder.offsets, weights.offsets)
        m.encoder.embeddings.copy_(weights.embeddings)
        for i, li in enumerate((0, 2, 4, 6)):
            m.pts_linears[li].weight.copy_(weights.pts_w[i]); m.pts_linears[li].bias.copy_(weights.pts_b[i])
            m.rgb_linears[li].weight.copy_(weights.rgb_w[i]); m.rgb_linears[li].bias.copy_(weights.rgb_b[i])
        m.geo_linear[0].weight.copy_(weights.geo_w); m.geo_linear[0].bias.copy_(weights.geo_b)
        m.output_linear[0].weight.copy_(weights.out_w); m.output_linear[0].bias.copy_(weights.out_b)
        nr = net.non_rigid_mlp.module if hasattr(net.non_rigid_mlp, "module") else net.non_rigid_mlp
        for i, li in enumerate(range(0, 14, 2)):
            nr.block_mlps[li].weight.copy_(weights.nr_w[i]); nr.block_mlps[li].bias.copy_(weights.nr_b[i])
    return net


@contextlib.contextmanager
def _injected_rand(t_rand):
    """`_stratified_sampling` draws torch.rand(z_vals.shape) (network.py:429); hand it ours instead."""
    if t_rand is None:
        yield
        return
    real = torch.rand

    def fake(*shape, **kw):
        shp = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
        assert shp == tuple(t_rand.shape), (shp, t_rand.shape)
        return t_rand.clone()

    torch.rand = fake
    try:
        yield
    finally:
        torch.rand = real


def reference_render_rays(net, frame, motion_weights_vol, *, iter_val, training, perturb, t_rand=None,
                          ignore_non_rigid=False):
    """Calls the reference's `_batchify_rays` -> `_render_rays` exactly as `Network.forward` does
    (network.py:573-615) with a given motion basis / weight volume."""
    _net_mod, cfg = load_reference()
    cfg.perturb = float(perturb)
    cfg.ignore_non_rigid_motions = bool(ignore_non_rigid)
    net.train(training)
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        emb_fn, _ = net.get_non_rigid_embedder(multires=cfg.non_rigid_motion_mlp.multires,
                                               is_identity=cfg.non_rigid_motion_mlp.i_embed, iter_val=iter_val)
    finally:
        os.chdir(cwd)
    posevec = frame.dst_posevec[None]
    nr_in = torch.zeros_like(posevec) * posevec if iter_val < cfg.non_rigid_motion_mlp.kick_in_iter else posevec
    packed = torch.cat([frame.rays_o, frame.rays_d, frame.near, frame.far], -1)
    with _injected_rand(t_rand):
        out = net._batchify_rays(
            packed, pos_embed_fn=net.pos_embed_fn, non_rigid_pos_embed_fn=emb_fn, non_rigid_mlp_input=nr_in,
            motion_scale_Rs=frame.motion_scale_Rs[None], motion_Ts=frame.motion_Ts[None],
            motion_weights_vol=motion_weights_vol, cnl_bbox_min_xyz=frame.cnl_bbox_min_xyz,
            cnl_bbox_scale_xyz=frame.cnl_bbox_scale_xyz, bgcolor=frame.bgcolor)
    return out
