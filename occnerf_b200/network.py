"""Host-side mirror of the reference's model interface for the per-ray rendering path.

`Network.forward / _render_rays / _query_mlp` keep the reference's signatures, argument meaning and return
dictionaries (core/nets/occnerf/network.py:542-549, 435-447, 164-192); parameter names and shapes follow
SURVEY.md appendix B so that reference checkpoints load with `load_state_dict`.  Behind them every stage is a
CUDA kernel of liboccnerf_b200.so (occnerf_b200/ops.py); there is no eager-PyTorch or CPU path.

Differences that are deliberate and visible to a caller:
  * `torch.rand` jitter (network.py:429) can be injected (`t_rand=`) so that runs are reproducible;
  * the in-place `point_counter[...] += 1` side effect (network.py:517) is applied by `apply_visibility()`
    from the returned `hits` so that data-parallel ranks can sum their votes first;
  * K-NN ids are int32 internally (the reference carries int64).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn as nn

from occnerf_b200 import mlp as M
from occnerf_b200 import ops

f32, i32 = torch.float32, torch.int32
TC_PASSES = {"tc1": 1, "tf32": 2, "tc3": 3, "tc3b1": 3}     # tcgen05 engines -> n_pass of csrc/mlp_tc.cu


# ----------------------------------------------------------------------------- parameter containers
def _xavier_(lin: nn.Linear, relu: bool):
    """`initseq` (core/utils/network_util.py:316-334)."""
    nn.init.xavier_uniform_(lin.weight, gain=math.sqrt(2.0) if relu else 1.0)
    nn.init.zeros_(lin.bias)


def _stack(dims):
    mods = []
    for i, o in zip(dims[:-1], dims[1:]):
        lin = nn.Linear(i, o)
        _xavier_(lin, True)
        mods += [lin, nn.ReLU(inplace=True)]
    return nn.ModuleList(mods)


class GridEncoder(nn.Module):
    """Same constructor, parameters (`embeddings`, `offsets`) and forward contract as the reference's
    gridencoder/grid.py:97-170 (hash grid type, align_corners=False)."""

    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, gridtype="hash", align_corners=False):
        super().__init__()
        if gridtype != "hash" or align_corners:
            raise ValueError("occnerf_b200.GridEncoder supports gridtype='hash', align_corners=False (what OccNeRF uses)")
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim, self.num_levels, self.level_dim = input_dim, num_levels, level_dim
        self.per_level_scale, self.base_resolution = per_level_scale, base_resolution
        self.log2_hashmap_size = log2_hashmap_size
        self.output_dim = num_levels * level_dim
        offsets, offset = [], 0
        self.max_params = 2 ** log2_hashmap_size
        for i in range(num_levels):
            resolution = int(np.ceil(base_resolution * per_level_scale ** i))
            params_in_level = min(self.max_params, (resolution + 1) ** input_dim)
            params_in_level = int(np.ceil(params_in_level / 8) * 8)
            offsets.append(offset)
            offset += params_in_level
        offsets.append(offset)
        self.register_buffer("offsets", torch.from_numpy(np.array(offsets, dtype=np.int32)))
        self.n_params = offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(offset, level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        self.embeddings.data.uniform_(-1e-4, 1e-4)

    def forward(self, inputs, bound=1):
        """gridencoder/grid.py:146-170: inputs in [-bound, bound] are mapped to [0, 1] (bound = 1 by default, a tensor bound
        broadcasts); pass bound=None for inputs that already are in [0, 1]."""
        if bound is not None:
            inputs = (inputs + bound) / (2 * bound)
        prefix = list(inputs.shape[:-1])
        out = ops.grid_encode(inputs.reshape(-1, self.input_dim), self.embeddings, self.offsets, self.per_level_scale,
                              self.base_resolution)
        return out.view(prefix + [self.output_dim])


class CanonicalMLP(nn.Module):
    """Parameter layout of canonical_mlps/occnerf_mlp.py:32-83 (mlp_depth=4, mlp_width=256, skips=[])."""

    def __init__(self, mlp_depth=4, mlp_width=256, bound=1.0, **_):
        super().__init__()
        assert mlp_depth == 4 and mlp_width == 256, "the fused kernels are built for the shipped 4x256 configuration"
        self.bound = bound
        self.encoder = GridEncoder(input_dim=4, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19,
                                   desired_resolution=2048 * bound)
        self.pts_linears = _stack([68, 256, 256, 256, 256])
        self.geo_linear = nn.Sequential(nn.Linear(256, 65))
        _xavier_(self.geo_linear[0], False)
        self.rgb_linears = _stack([131, 256, 256, 256, 256])
        self.output_linear = nn.Sequential(nn.Linear(256, 3))
        _xavier_(self.output_linear[0], False)

    def flat_params(self):
        out = []
        for i in (0, 2, 4, 6):
            out += [self.pts_linears[i].weight, self.pts_linears[i].bias]
        out += [self.geo_linear[0].weight, self.geo_linear[0].bias]
        for i in (0, 2, 4, 6):
            out += [self.rgb_linears[i].weight, self.rgb_linears[i].bias]
        out += [self.output_linear[0].weight, self.output_linear[0].bias]
        return out


class NonRigidMotionMLP(nn.Module):
    """Parameter layout of non_rigid_motion_mlps/mlp_offset.py:7-42."""

    def __init__(self):
        super().__init__()
        mods = []
        for di in (105, 128, 128, 128, 164, 128):
            lin = nn.Linear(di, 128)
            _xavier_(lin, True)
            mods += [lin, nn.ReLU()]
        last = nn.Linear(128, 3)
        last.weight.data.uniform_(-1e-5, 1e-5)
        last.bias.data.zero_()
        mods.append(last)
        self.block_mlps = nn.ModuleList(mods)

    def flat(self):
        lin = [self.block_mlps[i] for i in range(0, 14, 2)]
        return [l.weight for l in lin], [l.bias for l in lin]


class _Replicated(nn.Module):
    """Keeps the `.module.` level that nn.DataParallel puts into the reference's state_dict keys
    (network.py:68-72,142-146) without any of its behaviour."""

    def __init__(self, module):
        super().__init__()
        self.module = module


class HannEmbedder:
    """Stand-in for the closure `get_non_rigid_embedder` returns (hannw_fourier.py:48-62): carries the window
    weights; the encoding itself is evaluated inside the CUDA path."""

    def __init__(self, window):
        self.window = list(window)

    def __call__(self, xyz):
        return ops.hann_pe(xyz.reshape(-1, 3).contiguous(), self.window).view(*xyz.shape[:-1], 6 * len(self.window))


@dataclass
class RenderConfig:
    """The cfg values that parameterise the path."""
    N_samples: int = 128
    perturb: float = 1.0
    chunk: int = 32768
    netchunk_per_gpu: int = 300000
    total_bones: int = 24
    non_rigid_kick_in_iter: int = 100000
    non_rigid_full_band_iter: int = 200000
    non_rigid_multires: int = 6
    ignore_non_rigid_motions: bool = False
    bgcolor: tuple = (0.0, 0.0, 0.0)
    mlp_engine: str = "fp32"        # "fp32" (exact SIMT) | "tf32" (tcgen05 kind::tf32) | "tc3" (tcgen05 split-bf16) | "tc3b1" (tc3 forward, bf16 dgrad) | "tc1" (tcgen05 bf16)
    knn_mode: str = "grid"          # "grid" (per-cell candidate lists) | "tree" (3-level cluster tree) | "hier" | "brute"; same ids


# ----------------------------------------------------------------------------- differentiable stages
class _WarpFn(torch.autograd.Function):
    """(vol, Rs, Ts) -> (z, x_skel, mask); only mask carries a gradient, as in the reference: to the weight
    volume, and -- through F.grid_sample's grid gradient upstream, network.py:367-370 -- to motion_scale_Rs / motion_Ts
    (consumed only when the pose decoder trains, i.e. iter >= pose_decoder.kick_in_iter)."""

    @staticmethod
    def forward(ctx, vol, rays, t_rand, Rs, Ts, bmin, bscale, S, vol8):
        if vol8 is None:
            vol8 = ops.warp_pack_volume(vol.detach().contiguous().float(), Rs.shape[0])
        z, x_skel, mask = ops.warp_forward(rays, t_rand, Rs.detach(), Ts.detach(), vol.detach(), bmin, bscale, S, vol8=vol8)
        ctx.save_for_backward(rays, t_rand if t_rand is not None else torch.empty(0, device=rays.device), Rs.detach(), Ts.detach(), bmin, bscale, vol8)
        ctx.meta = (S, tuple(vol.shape), t_rand is not None)
        ctx.mark_non_differentiable(z, x_skel)
        return z, x_skel, mask

    @staticmethod
    def backward(ctx, _gz, _gx, g_mask):
        rays, t_rand, Rs, Ts, bmin, bscale, vol8 = ctx.saved_tensors
        S, vol_shape, has_rand = ctx.meta
        want_pose = ctx.needs_input_grad[3] or ctx.needs_input_grad[4]
        r = ops.warp_backward(rays, t_rand if has_rand else None, Rs, Ts, bmin, bscale, g_mask.contiguous(), S, vol_shape, vol8=vol8,
                              want_pose=want_pose)
        g_vol, g_Rs, g_Ts = r if want_pose else (r, None, None)
        return g_vol, None, None, g_Rs, g_Ts, None, None, None, None


class _CompositeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, mask, z, rays, bg, want_comp):
        raw_c, mask_c = raw.detach().contiguous(), mask.detach().contiguous()
        rgb, acc, depth, term, _, comp = ops.composite_forward(raw_c, mask_c, z, rays, bg, want_comp=want_comp)
        ctx.save_for_backward(raw_c, mask_c, z, rays, bg)
        ctx.want_comp = want_comp
        ctx.mark_non_differentiable(term)
        if comp is None:
            comp = torch.zeros(1, 1, device=raw.device)
        return rgb, acc, depth, term, comp

    @staticmethod
    def backward(ctx, g_rgb, g_acc, g_depth, _gt, g_comp):
        raw, mask, z, rays, bg = ctx.saved_tensors
        N = z.shape[0]
        dev = z.device
        g_rgb = g_rgb if g_rgb is not None else torch.zeros(N, 3, device=dev)
        g_acc = g_acc if g_acc is not None else torch.zeros(N, device=dev)
        g_depth = g_depth if g_depth is not None else torch.zeros(N, device=dev)
        g_raw, g_mask = ops.composite_backward(raw, mask, z, rays, bg, g_rgb.float(), g_acc.float(), g_depth.float(),
                                               g_comp.float() if (ctx.want_comp and g_comp is not None) else None)
        return g_raw, g_mask, None, None, None, None


class _QueryFn(torch.autograd.Function):
    """Canonical query of one chunk of (already offset) points with their multi-scale neighbour ids: surface geometry ->
    attention aggregation + hash encode -> MLP  (network.py:256-299 + occnerf_mlp.py:142-199).
    Differentiable w.r.t. the per-vertex feature table, the hash table and the 20 MLP tensors."""

    @staticmethod
    def forward(ctx, xyz, knn_idx, feats36, embeddings, net, shared, *mlp_params):
        m, dev = xyz.shape[0], xyz.device
        st = net._static()
        raw = torch.empty(m, 5, device=dev, dtype=f32)
        enc_in, _ = ops.sample_geometry(xyz, knn_idx, st["point_base"], st["point_norms"], net.bound, raw=raw)
        XB = torch.empty(m, M.XB_LD, device=dev, dtype=f32)
        feats_c = feats36.detach().contiguous()
        emb_c = embeddings.detach().contiguous()
        counter = net.point_counter.detach().contiguous().float()
        att_w = ops.aggregate_forward(knn_idx, counter, feats_c, XB.data_ptr() + 4 * M.X0_OFF, M.XB_LD,
                                      want_att=any(ctx.needs_input_grad))
        enc = net.cnl_mlp.module.encoder
        scales = ops.level_scales(float(np.log2(enc.per_level_scale)), enc.base_resolution, enc.num_levels, dev)
        ops.hashgrid_forward(enc_in, emb_c, enc.offsets, scales, out_ptr=XB.data_ptr() + 4 * M.H_OFF, ld=M.XB_LD,
                             run_length=ops.HASH_BWD_RUN)                                    # samples are ordered along rays
        W = M.MlpWeights(mlp_params)
        need_grad = any(ctx.needs_input_grad)
        engine = net._engine()
        saved = engine.forward(XB, raw, W, save=need_grad, shared=shared) if hasattr(engine, "bwd_pass") else \
            engine.forward(XB, raw, W, save=need_grad)
        if need_grad:
            ctx.state = dict(knn_idx=knn_idx, enc_in=enc_in, XB=XB, W=W, saved=saved, engine=engine, counter=counter,
                             offsets=enc.offsets, scales=scales, emb_shape=tuple(embeddings.shape), V=feats36.shape[0],
                             shared=shared, att_w=att_w)
            shared["pending"] = shared.get("pending", 0) + 1
        return raw

    @staticmethod
    def backward(ctx, g_raw):
        s = ctx.state
        if s is None:
            # the saved buffers (and the per-call gradient buffers the chunks share) are released by the first backward pass
            raise RuntimeError("occnerf_b200: the query node has already been differentiated once (retain_graph / a second backward "
                               "through the same forward is not supported: run the forward again)")
        g_raw = g_raw.contiguous().float()
        sh = s["shared"]
        last = sh["pending"] == 1
        if hasattr(s["engine"], "bwd_pass"):      # tensor-core engine: one weight-gradient buffer for all chunks of the call
            # (the weight-gradient kernel goes to a side stream and runs beside the scatter kernels below; joined by finish_wgrad)
            gXB, _ = s["engine"].backward(s["XB"], g_raw, s["W"], s["saved"], shared=sh, last=last, defer=True)
            g_params = [None] * 20
        else:
            gXB, g_params = s["engine"].backward(s["XB"], g_raw, s["W"], s["saved"])
        # the chunks of one _query_mlp call scatter into ONE table-gradient buffer (and one set of privatised vertex-gradient
        # replicas); the chunk whose backward runs last hands it to autograd, the others contribute None (= zero).  That
        # replaces a 59 MiB memset + a 59 MiB add per chunk by one memset per call.
        # With `Network.emb_grad_out` bound (a persistent, caller-zeroed buffer -- under data parallelism a view of the symmetric
        # all-reduce buffer) the table gradient is scattered straight into it and autograd receives None for the embeddings.
        ext = sh.get("g_emb_ext")
        if "g_emb" not in sh:
            sh["g_emb"] = ext if ext is not None else torch.zeros(s["emb_shape"], device=g_raw.device, dtype=f32)
            sh["g_priv"] = torch.zeros(ops.AGG_BWD_COPIES, s["V"], 36, device=g_raw.device, dtype=f32)
        ops.hashgrid_backward(gXB.data_ptr() + 4 * M.H_OFF, M.XB_LD, 0, s["enc_in"], s["offsets"], s["scales"], sh["g_emb"],
                              s["emb_shape"][1], run_length=ops.HASH_BWD_RUN)           # samples are ordered along rays
        ops.aggregate_backward(s["knn_idx"], s["counter"], gXB.data_ptr() + 4 * M.X0_OFF, M.XB_LD, s["V"], g_priv=sh["g_priv"],
                               att_w=s["att_w"])                                         # samples are ordered along rays
        if last and hasattr(s["engine"], "bwd_pass"):
            g_params = s["engine"].finish_wgrad(sh)
        ctx.state = None
        sh["pending"] -= 1
        g_emb = g_feats = None
        if sh["pending"] == 0:
            g_emb, g_feats = sh.pop("g_emb"), sh.pop("g_priv").sum(0)
            if ext is not None:
                g_emb = None
        return (None, None, g_feats, g_emb, None, None, *g_params)


class _VertexFn(torch.autograd.Function):
    """Per-vertex block (network.py:263-284 + occnerf_mlp.py:171-175): (point_dist, embeddings) -> feats36 (V,36) =
    (hash-grid features of the vertex's surface projection (32), point_cloud (3), 0).  One kernel each way for the
    geometry (csrc/vertex.cu) around the hash-grid kernels."""

    @staticmethod
    def forward(ctx, point_dist, embeddings, net):
        st = net._static()
        base, norms = st["point_base"], st["point_norms"]
        V, dev = base.shape[0], base.device
        pd = point_dist.detach().reshape(-1).contiguous().float()
        pc = base + pd[:, None]
        kidx = ops.knn(pc.contiguous(), st["base4"], [0, V], 3)[:, 0].contiguous()
        v_in = torch.empty(V, 4, device=dev, dtype=f32)
        feats36 = torch.empty(V, 36, device=dev, dtype=f32)
        ops.vertex_block_forward(base, pd, norms, kidx, net.bound, v_in, feats36.data_ptr() + 4 * 32, 36)
        enc = net.cnl_mlp.module.encoder
        scales = ops.level_scales(float(np.log2(enc.per_level_scale)), enc.base_resolution, enc.num_levels, dev)
        emb_c = embeddings.detach().contiguous()
        _, dy_dx, _, _ = ops.hashgrid_forward(v_in, emb_c, enc.offsets, scales, out_ptr=feats36.data_ptr(), ld=36, want_dy_dx=True)
        ctx.state = dict(pd=pd, kidx=kidx, v_in=v_in, dy_dx=dy_dx, scales=scales, offsets=enc.offsets, net=net,
                         emb_shape=tuple(embeddings.shape))
        ctx.mark_non_differentiable(pc)
        return feats36, pc

    @staticmethod
    def backward(ctx, g_feats, _g_pc):
        s = ctx.state
        st = s["net"]._static()
        g = g_feats.contiguous().float()
        V = g.shape[0]
        L, Cc = s["offsets"].shape[0] - 1, s["emb_shape"][1]
        ext = getattr(s["net"], "emb_grad_out", None)
        g_emb = ext if ext is not None else torch.zeros(s["emb_shape"], device=g.device, dtype=f32)
        ops.hashgrid_backward(g.data_ptr(), 36, 0, s["v_in"], s["offsets"], s["scales"], g_emb, Cc)
        if ext is not None:
            g_emb = None
        g_v_in = ops.hashgrid_input_backward(g.data_ptr(), 36, 0, s["dy_dx"], V, 4, Cc, L)
        g_pd = ops.vertex_block_backward(st["point_base"], s["pd"], st["point_norms"], s["kidx"], s["net"].bound, g_v_in,
                                         g.data_ptr() + 4 * 32, 36)
        ctx.state = None
        return g_pd[:, None], g_emb, None


# ----------------------------------------------------------------------------- the model
class Network(nn.Module):
    def __init__(self, cfg: RenderConfig | None = None):
        super().__init__()
        self.cfg = cfg or RenderConfig()
        self.non_rigid_mlp = _Replicated(NonRigidMotionMLP())
        self._cache = None
        # set by generate_neural_points()
        self.bound = None
        self.fps_index = None
        # optional per-frame prologue (pose refinement, motion basis, weight-volume decoder):
        # callable(dst_Rs, dst_Ts, cnl_gtfms, motion_weights_priors, dst_posevec, iter_val) -> (Rs, Ts, vol)
        self.prologue = None

    def install_prologue(self, pose_kick_in_iter=2000000):
        """Adds the reference's per-frame modules under their reference names (`mweight_vol_decoder`,
        `pose_decoder`, `motion_basis_computer`; state_dict-compatible) and wires `forward` to them."""
        from occnerf_b200 import prologue as P
        self.motion_basis_computer = P.MotionBasisComputer()
        self.mweight_vol_decoder = P.MotionWeightVolumeDecoder()
        self.pose_decoder = P.BodyPoseRefiner()
        dev = next(self.parameters()).device
        for m in (self.mweight_vol_decoder, self.pose_decoder):
            m.to(dev)

        self._pose_kick_in_iter = pose_kick_in_iter
        self.prologue = "modules"          # (not a closure: the modules are looked up on `self`, so copies of the network work)
        return self

    def _run_prologue(self, dst_Rs, dst_Ts, cnl_gtfms, priors, dst_posevec, iter_val):
        dst_Rs = dst_Rs[None]
        if iter_val >= self._pose_kick_in_iter:
            dst_Rs = self.pose_decoder.refine(dst_Rs, dst_posevec[None])
        Rs, Ts = self.motion_basis_computer(dst_Rs, dst_Ts[None], cnl_gtfms[None])
        return Rs, Ts, self.mweight_vol_decoder(motion_weights_priors=priors[None])[0]

    # -- per-subject state (network.py:90-146)
    def generate_neural_points(self, vertices, normals, fps_index, bbox_min=None, bbox_max=None, point_dist=None,
                               point_counter=None):
        """`vertices`/`normals` are the T-pose SMPL vertices and their normals (the reference derives them from the
        SMPL pkl + trimesh, network.py:92-98); `fps_index` the three farthest-point subsets (network.py:113-118)."""
        verts = torch.as_tensor(vertices, dtype=f32)
        V = verts.shape[0]
        if bbox_min is None:
            bbox_min, bbox_max = verts.min(0)[0] - 0.3, verts.max(0)[0] + 0.3
        self.bound = float(torch.max(torch.abs(torch.cat([torch.as_tensor(bbox_min).reshape(-1), torch.as_tensor(bbox_max).reshape(-1)]))))
        self.point_base = nn.Parameter(verts.clone(), requires_grad=False)
        pd = point_dist if point_dist is not None else (torch.rand(V, 1) * 2 - 1) * 1e-4
        self.point_dist = nn.Parameter(torch.as_tensor(pd, dtype=f32).clone(), requires_grad=True)
        pcn = point_counter if point_counter is not None else torch.ones(V)
        self.point_counter = nn.Parameter(torch.as_tensor(pcn, dtype=f32).clone(), requires_grad=False)
        self.register_buffer("point_norms", torch.as_tensor(normals, dtype=f32).clone(), persistent=False)
        self.fps_index = [torch.as_tensor(f, dtype=torch.int64).clone() for f in fps_index]
        self.cnl_mlp = _Replicated(CanonicalMLP(bound=self.bound))
        self._cache = None
        return self

    def deploy_mlps_to_secondary_gpus(self):
        """network.py:149-154: the reference moves the two MLPs to `cfg.secondary_gpus[0]` for nn.DataParallel.  Here every
        process owns one GPU and the whole model lives on it, so there is nothing to move; the method stays because every
        reference caller chains it (`model.cuda().deploy_mlps_to_secondary_gpus()`: run.py:37, eval.py:52, trainer.py:52)."""
        return self

    @property
    def point_cloud(self):
        return self.point_base + self.point_dist

    def _static(self):
        """Device-resident support arrays for the KNN: rebuilt when the module moves or when `point_base` / `point_norms`
        are overwritten (e.g. `load_state_dict` of a reference checkpoint copies into them in place)."""
        dev = self.point_base.device
        key = (dev, self.point_base.data_ptr(), self.point_base._version, self.point_norms.data_ptr(), self.point_norms._version)
        if self._cache is None or self._cache["key"] != key:
            base = self.point_base.detach()
            fps = [f.to(dev) for f in self.fps_index]
            sup = torch.cat([base] + [base[f] for f in fps], 0)
            gid = torch.cat([torch.arange(base.shape[0], device=dev)] + fps).to(i32)
            lb = np.cumsum([0, base.shape[0]] + [int(f.shape[0]) for f in fps]).tolist()
            self._cache = dict(device=dev, key=key, supports4=ops.to_float4(sup), support_gid=gid.contiguous(), level_begin=lb,
                               base4=ops.to_float4(base), point_base=base.contiguous().float(),
                               point_norms=self.point_norms.contiguous().float(), fps=fps)
        return self._cache

    def _knn_structure(self, name):
        """Acceleration structures of the exact searches, built on first use (each is a few hundred library launches once per
        subject): 'grid' = per-cell candidate lists, 'tree' = 3-level cluster tree, 'hier' = the two 2-level hierarchies."""
        st = self._static()
        if name not in st:
            base, fps = st["point_base"], st["fps"]
            if name == "grid":
                st["grid"] = ops.build_knn_grid(base, fps)
            elif name == "tree":
                st["tree"] = ops.build_knn_tree(base, fps)
            else:       # cluster hierarchies for the pruned exact search: level 0 around level 2, level 1 around level 3
                st["hier"] = dict(hier0=ops.build_knn_hierarchy(base, base[fps[1]]), gid2=fps[1].to(i32).contiguous(),
                                  hier1=ops.build_knn_hierarchy(base[fps[0]], base[fps[2]]), gid1=fps[0].to(i32).contiguous(),
                                  gid3=fps[2].to(i32).contiguous())
        return st[name]

    def _engine(self):
        e = self.cfg.mlp_engine
        if e == "fp32":
            return M.MlpSimt()
        from occnerf_b200 import mlp_tc
        if e == "tc3b1":         # split-bf16 forward (fp32-grade outputs), bf16-operand data gradients
            return mlp_tc.MlpTc(n_pass=3, bwd_pass=1)
        if e not in TC_PASSES:
            raise ValueError(f"unknown mlp_engine {e!r}")
        return mlp_tc.MlpTc(n_pass=TC_PASSES[e])

    # -- per-vertex block (network.py:263-284 + occnerf_mlp.py:171-175), once per call instead of once per chunk
    def vertex_features(self):
        """-> (feats36 (V,36), point_cloud (V,3)); differentiable w.r.t. point_dist and the hash table."""
        return _VertexFn.apply(self.point_dist, self.cnl_mlp.module.encoder.embeddings, self)

    # -- reference API: network.py:164-192
    def _query_mlp(self, pos_xyz, rays_d, pos_embed_fn, non_rigid_pos_embed_fn, non_rigid_mlp_input, _feats36=None,
                   _return_knn=False):
        pos_flat = pos_xyz.reshape(-1, pos_xyz.shape[-1])
        self._group_stride = pos_xyz.shape[-2] if pos_xyz.dim() >= 3 else 1     # samples per ray (warp = 32 rays at one depth)
        feats36 = _feats36 if _feats36 is not None else self.vertex_features()[0]
        window, cond = None, None
        if not self.cfg.ignore_non_rigid_motions:
            if not hasattr(non_rigid_pos_embed_fn, "window"):
                raise TypeError("non_rigid_pos_embed_fn must be the HannEmbedder returned by get_non_rigid_embedder()")
            window = non_rigid_pos_embed_fn.window
            # the condition tensor is never inspected on the device (that would be a host sync per call and cannot be
            # captured in a CUDA graph): `forward` passes None before kick_in_iter, the reference passes zeros
            # (network.py:579-583); both give the same offsets
            if non_rigid_mlp_input is not None:
                cond = non_rigid_mlp_input.reshape(1, -1).float()
        # closed Hann window (all of training before kick_in_iter): every sample sees the same MLP input [cond, 0], so the
        # offset is one constant 3-vector, evaluated once per call instead of once per chunk
        self._nr_const = None
        if window is not None and all(v == 0.0 for v in window):
            nw, nb = self.non_rigid_mlp.module.flat()
            if cond is None:
                # ... and only when the non-rigid MLP's tensors have changed (they get no gradient before kick_in_iter)
                key = tuple((t.data_ptr(), t._version) for t in nw + nb)
                if getattr(self, "_nr_const_key", None) != key:
                    self._nr_const_val = M.nonrigid_offsets(pos_flat[:1].detach().contiguous().float(), None, window, nw, nb, return_const=True)
                    self._nr_const_key = key
                self._nr_const = self._nr_const_val
            else:                       # a caller-supplied condition code may change between calls: one-row evaluation per call
                self._nr_const = M.nonrigid_offsets(pos_flat[:1].detach().contiguous().float(), cond, window, nw, nb, return_const=True)
        # non-rigid offsets (network.py:225-232) and the multi-scale neighbour search (network.py:236-255) run under no_grad
        # in the reference and do not depend on the chunking: one search over all points keeps the SMs full (per-chunk
        # launches of ~2 waves lose a third of their time to the tail), the memory-heavy part below stays chunked
        chunk = self.cfg.netchunk_per_gpu
        xyz_all = pos_flat.detach().contiguous().float()
        if window is not None:
            if self._nr_const is not None:
                xyz_all = xyz_all + self._nr_const
            else:
                nw, nb = self.non_rigid_mlp.module.flat()
                moved = torch.empty_like(xyz_all)
                n_pass = TC_PASSES.get(self.cfg.mlp_engine)
                if n_pass is not None:           # tensor-core engines: the fused tcgen05 chain (exact fp32 mode: SIMT GEMMs)
                    packed = ops.nonrigid_pack(nw, nb, cond, n_pass)
                    for i in range(0, xyz_all.shape[0], chunk):
                        ops.nonrigid_forward_tc(xyz_all[i:i + chunk], window, packed, n_pass, out=moved[i:i + chunk])
                else:
                    for i in range(0, xyz_all.shape[0], chunk):
                        moved[i:i + chunk] = M.nonrigid_offsets(xyz_all[i:i + chunk], cond, window, nw, nb)
                xyz_all = moved
        knn_all = self._knn(xyz_all, self._group_stride)
        cm = self.cnl_mlp.module
        raws, shared = [], {"g_emb_ext": getattr(self, "emb_grad_out", None)}
        for i in range(0, pos_flat.shape[0], chunk):
            raws.append(_QueryFn.apply(xyz_all[i:i + chunk], knn_all[i:i + chunk], feats36, cm.encoder.embeddings, self, shared,
                                       *cm.flat_params()))
        raws_flat = raws[0] if len(raws) == 1 else torch.cat(raws, 0)
        out = {"raws": raws_flat.reshape(list(pos_xyz.shape[:-1]) + [5])}
        if _return_knn:
            out["knn_idxs"] = knn_all
        return out

    def _knn(self, xyz, group_stride):
        """(m,3) -> (m,4,10) int32 vertex ids on the four levels (exact; the three modes return identical ids)."""
        st = self._static()
        gs = max(1, int(group_stride))
        if self.cfg.knn_mode == "brute":
            return ops.knn(xyz, st["supports4"], st["level_begin"], 10, support_gid=st["support_gid"])
        if self.cfg.knn_mode == "grid":
            return ops.knn_grid(xyz, gs, self._knn_structure("grid"))
        if self.cfg.knn_mode == "tree":
            return ops.knn_tree(xyz, gs, self._knn_structure("tree"))
        h = self._knn_structure("hier")
        knn_idx = torch.empty(xyz.shape[0], 4, 10, device=xyz.device, dtype=i32)
        ops.knn_hier(xyz, gs, *h["hier0"], knn_idx, 0, 2, None, h["gid2"])
        ops.knn_hier(xyz, gs, *h["hier1"], knn_idx, 1, 3, h["gid1"], h["gid3"])
        return knn_idx

    def get_non_rigid_embedder(self, multires, is_identity, iter_val):
        return HannEmbedder(ops.hann_window(iter_val, self.cfg.non_rigid_kick_in_iter, self.cfg.non_rigid_full_band_iter,
                                            multires)), 6 * multires

    # -- reference API: network.py:435-525
    def _render_rays(self, ray_batch, motion_scale_Rs, motion_Ts, motion_weights_vol, cnl_bbox_min_xyz,
                     cnl_bbox_scale_xyz, pos_embed_fn, non_rigid_pos_embed_fn, non_rigid_mlp_input=None, bgcolor=None,
                     t_rand=None, _feats=None, _vol8=None, **_):
        cfg = self.cfg
        S = cfg.N_samples
        rays = ray_batch.contiguous().float()
        N = rays.shape[0]
        dev = rays.device
        if cfg.perturb > 0.0 and t_rand is None:
            t_rand = torch.rand(N, S, device=dev)
        if cfg.perturb <= 0.0:
            t_rand = None
        Rs = motion_scale_Rs.reshape(-1, 3, 3).contiguous().float()
        Ts = motion_Ts.reshape(-1, 3).contiguous().float()
        z, x_skel, mask = _WarpFn.apply(motion_weights_vol, rays, t_rand.contiguous() if t_rand is not None else None, Rs, Ts,
                                        cnl_bbox_min_xyz.contiguous().float(), cnl_bbox_scale_xyz.contiguous().float(), S, _vol8)
        feats36, pc = _feats if _feats is not None else self.vertex_features()
        q = self._query_mlp(pos_xyz=x_skel, rays_d=None, pos_embed_fn=pos_embed_fn,
                            non_rigid_pos_embed_fn=non_rigid_pos_embed_fn, non_rigid_mlp_input=non_rigid_mlp_input,
                            _feats36=feats36)
        raw = q["raws"]
        bg = (bgcolor if bgcolor is not None else torch.tensor(cfg.bgcolor)).to(dev).float().contiguous()
        rgb, acc, depth, term, comp = _CompositeFn.apply(raw, mask, z, rays, bg, self.training)
        out = {"rgb": rgb, "alpha": acc, "depth": depth, "comp_loss": comp}
        if self.training:
            out["hits"] = ops.visibility_hits(depth.detach(), term, x_skel, ops.to_float4(pc.detach()))
        return out

    def _batchify_rays(self, rays_flat, **kwargs):
        """network.py:307-317."""
        t_rand = kwargs.pop("t_rand", None)
        feats = self.vertex_features()
        # the per-frame weight volume in the corner-packed layout of K1 (csrc/warp.cu): once per frame, not once per ray chunk
        vol = kwargs["motion_weights_vol"]
        vol8 = ops.warp_pack_volume(vol.detach().contiguous().float(), self.cfg.total_bones) if vol.shape[0] >= self.cfg.total_bones else None
        all_ret = {}
        for i in range(0, rays_flat.shape[0], self.cfg.chunk):
            tr = t_rand[i:i + self.cfg.chunk] if t_rand is not None else None
            ret = self._render_rays(rays_flat[i:i + self.cfg.chunk], t_rand=tr, _feats=feats, _vol8=vol8, **kwargs)
            for k, v in ret.items():
                all_ret.setdefault(k, []).append(v)
        out = {}
        for k, v in all_ret.items():
            if k == "hits":
                # the reference increments point_counter once per ray chunk (network.py:502-517, inside _render_rays): the 0/1
                # votes of the chunks add up.  (It also lets later chunks see the updated counter; here all chunks of a call
                # see the counter as it was at the start -- documented difference, only for calls with more than cfg.chunk rays.)
                out[k] = torch.stack(v, 0).sum(0)
            elif k == "comp_loss" and not self.training:
                out[k] = torch.cat([c.reshape(-1) for c in v], 0) if len(v) > 1 else v[0]   # (n_chunks,) placeholders as upstream
            else:
                out[k] = v[0] if len(v) == 1 else torch.cat(v, 0)
        return out

    # -- optional persistent destination of the hash-table gradient
    def bind_emb_grad(self, buf=None):
        """From now on both hash-grid backward calls of a step (sample chunks, per-vertex block) scatter into ONE persistent buffer
        and autograd receives None for the embeddings: one 59 MiB memset per step instead of two and no 59 MiB accumulation add.
        `buf`: flat or table-shaped fp32 tensor of the embeddings' size (data parallel: SwitchReducer.table_view, so that the
        all-reduce finds the gradient in symmetric memory); None allocates one.  Call `zero_bound_grads()` before every backward
        and `attach_bound_grads()` after it."""
        emb = self.cnl_mlp.module.encoder.embeddings
        if buf is None:
            buf = torch.zeros(emb.numel(), device=emb.device, dtype=f32)
        if buf.numel() != emb.numel() or buf.dtype != f32 or not buf.is_contiguous() or buf.device != emb.device:
            raise RuntimeError("bind_emb_grad: the buffer must be a contiguous fp32 tensor of the embeddings' size on their device")
        self.emb_grad_out = buf

    def zero_bound_grads(self):
        if getattr(self, "emb_grad_out", None) is not None:
            self.emb_grad_out.zero_()

    def attach_bound_grads(self):
        if getattr(self, "emb_grad_out", None) is not None:
            emb = self.cnl_mlp.module.encoder.embeddings
            emb.grad = self.emb_grad_out.view_as(emb)

    def apply_visibility(self, hits):
        """The reference's in-place `point_counter[knn_index] += 1.` (network.py:517)."""
        with torch.no_grad():
            self.point_counter += hits.to(self.point_counter.dtype)

    # -- reference API: network.py:542-623
    def forward(self, rays, dst_Rs, dst_Ts, cnl_gtfms, motion_weights_priors, dst_posevec=None, near=None, far=None,
                iter_val=1e7, **kwargs):
        cfg = self.cfg
        if self.prologue is None:
            raise RuntimeError("Network.prologue is not set: the per-frame prologue (pose refinement, motion basis, "
                               "weight-volume decoder) is outside the ray path; install occnerf_b200.prologue.Prologue "
                               "or pass precomputed tensors to _batchify_rays()")
        run = self.prologue if callable(self.prologue) else self._run_prologue
        motion_scale_Rs, motion_Ts, vol = run(dst_Rs, dst_Ts, cnl_gtfms, motion_weights_priors, dst_posevec, iter_val)
        emb_fn, _ = self.get_non_rigid_embedder(cfg.non_rigid_multires, 0, iter_val)
        if iter_val < cfg.non_rigid_kick_in_iter:
            nr_in = None
        else:
            nr_in = dst_posevec[None, ...]
        rays_o, rays_d = rays
        rays_shape = rays_d.shape
        rays_o = rays_o.reshape(-1, 3).float()
        rays_d = rays_d.reshape(-1, 3).float()
        packed = torch.cat([rays_o, rays_d, near, far], -1)
        keep = {k: kwargs[k] for k in ("cnl_bbox_min_xyz", "cnl_bbox_scale_xyz", "bgcolor", "t_rand") if k in kwargs}
        all_ret = self._batchify_rays(packed, pos_embed_fn=None, non_rigid_pos_embed_fn=emb_fn, non_rigid_mlp_input=nr_in,
                                      motion_scale_Rs=motion_scale_Rs, motion_Ts=motion_Ts, motion_weights_vol=vol, **keep)
        for k in list(all_ret):
            if k == "comp_loss":
                all_ret[k] = all_ret[k].reshape(-1)
            elif k != "hits":
                all_ret[k] = all_ret[k].reshape(list(rays_shape[:-1]) + list(all_ret[k].shape[1:]))
        return all_ret
