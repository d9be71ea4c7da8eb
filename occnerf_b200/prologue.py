"""Per-frame prologue of `Network.forward` (network.py:558-597): pose refinement, motion basis, motion-weight
volume decoder.  It runs once per frame and produces the ray path's inputs (`motion_scale_Rs`, `motion_Ts`,
`motion_weights_vol`); SURVEY.md section 8(f) rank 1.  The modules keep the reference's names and parameters (checkpoints load).
On a CUDA device the small stages are native kernels (csrc/prologue.cu): the 24-bone motion basis (one launch instead of ~50),
the pose refiner MLP + Rodrigues (one launch), and the decoder's softmax(logits + log prior) with its gradient.  The decoder's five
ConvTranspose3d exist natively too -- tf32 tensor-core GEMMs on the reference weight layout with their data and weight gradients
(csrc/deconv.cu, 3.7 GMAC per pass at batch 1; `MotionWeightVolumeDecoder.native`, the default); the library form (cuDNN) remains
as the cross-check of the tests.
When the inputs of the first two require gradients (pose refinement training) they fall back to the differentiable torch form.

  MotionBasisComputer         core/utils/network_util.py:138-200   (FK chain evaluated level by level of the SMPL tree)
  MotionWeightVolumeDecoder   mweight_vol_decoders/deconv_vol_decoder.py:8-33 + network_util.py:12-50
  BodyPoseRefiner             pose_decoders/mlp_delta_body_pose.py:35-41 (kick_in_iter = 2e6 in every shipped yaml)
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

SMPL_PARENT = {1: 0, 2: 0, 3: 0, 4: 1, 5: 2, 6: 3, 7: 4, 8: 5, 9: 6, 10: 7, 11: 8, 12: 9, 13: 9, 14: 9, 15: 12, 16: 13,
               17: 14, 18: 16, 19: 17, 20: 18, 21: 19, 22: 20, 23: 21}


def _tree_levels():
    depth = {0: 0}
    for i in range(1, 24):
        depth[i] = depth[SMPL_PARENT[i]] + 1
    levels = []
    for d in range(1, max(depth.values()) + 1):
        idx = [i for i in range(24) if depth[i] == d]
        levels.append((idx, [SMPL_PARENT[i] for i in idx]))
    return levels


_LEVELS = _tree_levels()
_LEVEL_IDX = {}


def _init_seq(seq):
    """xavier-uniform with the gain of the following activation, zero bias (network_util.py:265-334)."""
    mods = list(seq)
    for i, m in enumerate(mods):
        if isinstance(m, (nn.Linear, nn.ConvTranspose3d)):
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            gain = nn.init.calculate_gain("leaky_relu", 0.2) if isinstance(nxt, nn.LeakyReLU) else \
                (math.sqrt(2.0) if isinstance(nxt, nn.ReLU) else 1.0)
            nn.init.xavier_uniform_(m.weight, gain=gain)
            nn.init.zeros_(m.bias)


def _affine_inverse(T):
    """Inverse of a batch of affine 4x4 matrices [A t; 0 1] by cofactors (network_util.py:190 calls torch.inverse, whose
    LU path synchronises with the host and cannot be captured in a CUDA graph; same result to fp32 rounding)."""
    A, t = T[:, :3, :3], T[:, :3, 3]
    a, b, c = A[:, :, 0], A[:, :, 1], A[:, :, 2]
    bc, ca, ab = torch.cross(b, c, dim=1), torch.cross(c, a, dim=1), torch.cross(a, b, dim=1)
    det = (a * bc).sum(1, keepdim=True)
    Ainv = torch.stack([bc, ca, ab], 1) / det[:, :, None]
    out = torch.zeros_like(T)
    out[:, :3, :3] = Ainv
    out[:, :3, 3] = -torch.matmul(Ainv, t[:, :, None])[:, :, 0]
    out[:, 3, 3] = 1.0
    return out


class _VolumeSoftmax(torch.autograd.Function):
    """softmax_channels(logits + log prior) per voxel (occnerf_weight_volume_forward / _backward)."""

    @staticmethod
    def forward(ctx, logits, priors):
        from occnerf_b200 import ops
        vol = ops.weight_volume_forward(logits.detach().contiguous().float(), priors.contiguous().float())
        ctx.save_for_backward(vol)
        return vol

    @staticmethod
    def backward(ctx, g):
        from occnerf_b200 import ops
        (vol,) = ctx.saved_tensors
        return ops.weight_volume_backward(vol, g.contiguous().float()), None


class _DecoderFn(torch.autograd.Function):
    """MotionWeightVolumeDecoder on its native kernels (csrc/deconv.cu + the softmax of csrc/prologue.cu), batch 1:
    const_embedding -> Linear(256, 1024) -> LeakyReLU -> 5 x ConvTranspose3d(4, 2, 1) (LeakyReLU between) -> softmax(. + log prior).
    args: embedding [256], priors [25, 32, 32, 32], lin_w, lin_b, then (w, b) of the five transposed convolutions.
    Pre-activations are kept for the backward pass; `exact` follows torch.backends.cudnn.allow_tf32 (False -> 3 x tf32)."""

    @staticmethod
    def forward(ctx, emb, priors, *wb):
        from occnerf_b200 import ops
        exact = not torch.backends.cudnn.allow_tf32
        wb = [t.detach().contiguous().float() for t in wb]
        e = emb.detach().contiguous().float()
        ys = [ops.decoder_linear_forward(wb[0], wb[1], e)]                   # [1024] = [1024, 1^3]
        D = 1
        for l in range(5):
            ys.append(ops.deconv3d_forward(wb[2 + 2 * l], wb[3 + 2 * l], ys[-1], D, 0.2, exact))
            D *= 2
        logits = ys.pop().view(-1, D, D, D)
        vol = ops.weight_volume_forward(logits, priors.contiguous().float())
        ctx.save_for_backward(e, vol, *ys, *wb)
        ctx.exact = exact
        ctx.grad_out = getattr(_DecoderFn, "grad_out", None)
        return vol

    @staticmethod
    def backward(ctx, g_vol):
        from occnerf_b200 import ops
        e, vol, *rest = ctx.saved_tensors
        ys, wb = rest[:5], rest[5:]
        g = ops.weight_volume_backward(vol, g_vol.contiguous().float()).view(vol.shape[0], -1)
        grads = [None] * 12
        D = 16
        for l in range(4, -1, -1):
            dst = ctx.grad_out[2 + 2 * l] if ctx.grad_out is not None else None
            if dst is not None:
                dst = dst.view_as(dst)          # a fresh tensor object on the same memory: autograd adopts it instead of cloning it
            dW, db, g = ops.deconv3d_backward(wb[2 + 2 * l], ys[l], g, D, 0.2, ctx.exact, need_dyin=True, dW_out=dst)
            grads[2 + 2 * l], grads[3 + 2 * l] = dW, db
            D //= 2
        grads[0], grads[1], de = ops.decoder_linear_backward(wb[0], e, g.view(-1))
        return (de, None, *grads)


class MotionBasisComputer(nn.Module):
    def forward(self, dst_Rs, dst_Ts, cnl_gtfms):
        if dst_Rs.is_cuda and dst_Rs.shape[0] == 1 and not (dst_Rs.requires_grad or dst_Ts.requires_grad):
            from occnerf_b200 import ops
            Rs, Ts = ops.motion_basis(dst_Rs[0].contiguous().float(), dst_Ts[0].contiguous().float(), cnl_gtfms[0].contiguous().float())
            return Rs[None], Ts[None]
        B = dst_Rs.shape[0]
        G = torch.zeros(B, 24, 4, 4, dtype=dst_Rs.dtype, device=dst_Rs.device)
        G[:, :, :3, :3] = dst_Rs
        G[:, :, :3, 3] = dst_Ts
        G[:, :, 3, 3] = 1.0
        glob = [None] * 24
        glob[0] = G[:, 0]
        key = str(G.device)
        if key not in _LEVEL_IDX:       # device-resident index tensors (a Python list index would be an H2D copy per call)
            _LEVEL_IDX[key] = [torch.tensor(idx, device=G.device) for idx, _ in _LEVELS]
        for (idx, par), idx_t in zip(_LEVELS, _LEVEL_IDX[key]):
            prod = torch.matmul(torch.stack([glob[p] for p in par], 1), G.index_select(1, idx_t))
            for j, i in enumerate(idx):
                glob[i] = prod[:, j]
        dst = torch.stack(glob, 1).view(-1, 4, 4)
        f = torch.matmul(cnl_gtfms.view(-1, 4, 4), _affine_inverse(dst)).view(B, 24, 4, 4)
        return f[:, :, :3, :3], f[:, :, :3, 3]


class ConvDecoder3D(nn.Module):
    def __init__(self, embedding_size=256, volume_size=32, voxel_channels=25):
        super().__init__()
        self.block_mlp = nn.Sequential(nn.Linear(embedding_size, 1024), nn.LeakyReLU(0.2))
        block_conv, inc, outc = [], 1024, 512
        for _ in range(int(np.log2(volume_size)) - 1):
            block_conv += [nn.ConvTranspose3d(inc, outc, 4, 2, 1), nn.LeakyReLU(0.2)]
            if inc == outc:
                outc = inc // 2
            else:
                inc = outc
        block_conv.append(nn.ConvTranspose3d(inc, voxel_channels, 4, 2, 1))
        self.block_conv = nn.Sequential(*block_conv)
        _init_seq(self.block_mlp)
        _init_seq(self.block_conv)

    def forward(self, embedding):
        return self.block_conv(self.block_mlp(embedding).view(-1, 1024, 1, 1, 1))


class MotionWeightVolumeDecoder(nn.Module):
    def __init__(self, embedding_size=256, volume_size=32, total_bones=24):
        super().__init__()
        self.const_embedding = nn.Parameter(torch.randn(embedding_size))
        self.decoder = ConvDecoder3D(embedding_size, volume_size, total_bones + 1)

    # True (default): the five transposed convolutions on the native kernels of csrc/deconv.cu (forward, data and weight gradients;
    # parity-tested in tests/test_deconv_gpu.py).  B200, forward + backward per step, tf32: 0.96 ms against 1.09 ms for the library
    # form (cuDNN incl. its layout passes; gpurun_out/r2t_decoder_bench.json).  OCCNERF_NATIVE_DECODER=0 keeps the library form
    # (the cross-check of the tests).
    native = os.environ.get("OCCNERF_NATIVE_DECODER", "1") != "0"

    def forward(self, motion_weights_priors, **_):
        if self.native and self.const_embedding.is_cuda and motion_weights_priors.shape[0] == 1:
            d = self.decoder
            convs = [m for m in d.block_conv if isinstance(m, nn.ConvTranspose3d)]
            wb = [d.block_mlp[0].weight, d.block_mlp[0].bias] + [t for c in convs for t in (c.weight, c.bias)]
            return _DecoderFn.apply(self.const_embedding, motion_weights_priors[0], *wb)[None]
        logits = self.decoder(self.const_embedding[None])                  # library (cuDNN) transposed convolutions
        if logits.is_cuda and logits.shape[0] == 1:
            return _VolumeSoftmax.apply(logits[0], motion_weights_priors[0])[None]
        return F.softmax(logits + torch.log(motion_weights_priors), dim=1)


def rodrigues(rvec):
    """network_util.py:98-127."""
    theta = torch.sqrt(1e-5 + torch.sum(rvec ** 2, dim=1))
    r = rvec / theta[:, None]
    c, s = torch.cos(theta), torch.sin(theta)
    x, y, z = r[:, 0], r[:, 1], r[:, 2]
    return torch.stack((x ** 2 + (1. - x ** 2) * c, x * y * (1. - c) - z * s, x * z * (1. - c) + y * s,
                        x * y * (1. - c) + z * s, y ** 2 + (1. - y ** 2) * c, y * z * (1. - c) - x * s,
                        x * z * (1. - c) - y * s, y * z * (1. - c) + x * s, z ** 2 + (1. - z ** 2) * c), dim=1).view(-1, 3, 3)


class BodyPoseRefiner(nn.Module):
    def __init__(self, embedding_size=69, mlp_width=256, mlp_depth=4, total_bones=24):
        super().__init__()
        mods = [nn.Linear(embedding_size, mlp_width), nn.ReLU()]
        for _ in range(mlp_depth - 1):
            mods += [nn.Linear(mlp_width, mlp_width), nn.ReLU()]
        self.total_bones = total_bones - 1
        mods.append(nn.Linear(mlp_width, 3 * self.total_bones))
        self.block_mlps = nn.Sequential(*mods)
        _init_seq(self.block_mlps)
        self.block_mlps[-1].weight.data.uniform_(-1e-5, 1e-5)
        self.block_mlps[-1].bias.data.zero_()

    def forward(self, pose_input):
        return {"Rs": rodrigues(self.block_mlps(pose_input).view(-1, 3)).view(-1, self.total_bones, 3, 3)}

    def refine(self, dst_Rs, pose_input):
        """dst_Rs (1,24,3,3), pose_input (1,69) -> dst_Rs with the non-root rotations multiplied by the predicted corrections
        (network.py:558-570); one native launch when nothing here needs a gradient, the torch form otherwise."""
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if dst_Rs.is_cuda and dst_Rs.shape[0] == 1 and not needs_grad:
            from occnerf_b200 import ops
            lin = [m for m in self.block_mlps if isinstance(m, nn.Linear)]
            return ops.pose_refine([l.weight for l in lin], [l.bias for l in lin], pose_input.reshape(-1).contiguous().float(),
                                   dst_Rs[0].contiguous().float())[None]
        refined = self.forward(pose_input)["Rs"]
        no_root = torch.matmul(dst_Rs[:, 1:].reshape(-1, 3, 3), refined.reshape(-1, 3, 3)).reshape(-1, 23, 3, 3)
        return torch.cat([dst_Rs[:, 0:1], no_root], dim=1)


class Prologue(nn.Module):
    """callable(dst_Rs, dst_Ts, cnl_gtfms, motion_weights_priors, dst_posevec, iter_val) -> (Rs, Ts, vol)
    with the reference's control flow (network.py:558-597): pose refinement from `pose_kick_in_iter` on
    (2 000 000 in every shipped yaml, i.e. never while training, always at render time where iter_val = 1e7)."""

    def __init__(self, pose_kick_in_iter=2000000):
        super().__init__()
        self.motion_basis_computer = MotionBasisComputer()
        self.mweight_vol_decoder = MotionWeightVolumeDecoder()
        self.pose_decoder = BodyPoseRefiner()
        self.pose_kick_in_iter = pose_kick_in_iter

    def forward(self, dst_Rs, dst_Ts, cnl_gtfms, motion_weights_priors, dst_posevec, iter_val):
        dst_Rs = dst_Rs[None]
        if iter_val >= self.pose_kick_in_iter:
            dst_Rs = self.pose_decoder.refine(dst_Rs, dst_posevec[None])
        Rs, Ts = self.motion_basis_computer(dst_Rs, dst_Ts[None], cnl_gtfms[None])
        vol = self.mweight_vol_decoder(motion_weights_priors=motion_weights_priors[None])[0]
        return Rs, Ts, vol
