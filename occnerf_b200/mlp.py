"""Host-side drivers of the canonical MLP (occnerf_mlp.py:183-199) and the non-rigid offset MLP
(mlp_offset.py:45-62) on top of the C ABI.

Two engines share one interface:
  * MlpSimt  -- exact fp32 (occnerf_sgemm), forward + backward.  The "precise" mode.
  * MlpTc    -- fused tcgen05/TMEM kernel (occnerf_mlp_forward_tc), see csrc/mlp_tc.cu.

Sample-major activation buffer shared by both (no concatenation copies anywhere):
  XB [M,132] : 0..63 geo features (written by the geo layer) | 64..98 agg | 99 var | 100..131 hash features
so that the geometry trunk reads XB[:,64:132] (K=68) and the colour trunk reads XB (K=132, with a zero
weight column inserted at 99 for the variance slot, which the reference's colour trunk does not see).
"""
from __future__ import annotations

import ctypes as C

import torch

from occnerf_b200 import _lib, ops
from occnerf_b200._lib import call, ptr, stream

f32 = torch.float32
XB_LD = 132
X0_OFF = 64          # XB[:, 64:132] = (agg35, var1, h32)
H_OFF = 100          # XB[:, 100:132] = hash features
FLOP_FWD = 923136.0  # per sample, SURVEY.md section 8d


class MlpWeights:
    """Flat view of the 20 canonical-MLP tensors in the order the autograd.Function receives them."""
    ORDER = ["pts_w0", "pts_b0", "pts_w1", "pts_b1", "pts_w2", "pts_b2", "pts_w3", "pts_b3", "geo_w", "geo_b",
             "rgb_w0", "rgb_b0", "rgb_w1", "rgb_b1", "rgb_w2", "rgb_b2", "rgb_w3", "rgb_b3", "out_w", "out_b"]

    def __init__(self, tensors):
        assert len(tensors) == 20
        t = [x.detach().contiguous() for x in tensors]
        self.pts_w, self.pts_b = t[0:8:2], t[1:8:2]
        self.geo_w, self.geo_b = t[8], t[9]
        self.rgb_w, self.rgb_b = t[10:18:2], t[11:18:2]
        self.out_w, self.out_b = t[18], t[19]
        assert self.pts_w[0].shape == (256, 68) and self.rgb_w[0].shape == (256, 131) and self.geo_w.shape == (65, 256)


def _pad_rgb0(w):
    """(256,131) -> (256,132): zero column at the variance slot."""
    z = torch.zeros(w.shape[0], 1, device=w.device, dtype=w.dtype)
    return torch.cat([w[:, :99], z, w[:, 99:]], 1).contiguous()


def _splits(M):
    return max(1, min(256, (M + 4095) // 4096))


def _gemm(A, sAi, sAr, B, sBr, sBj, Cp, ldc, Mi, Nj, Kr, bias=None, mask=None, ldmask=0, relu=False, accum=False,
          split_k=1):
    flags = (_lib.GEMM_BIAS if bias else 0) | (_lib.GEMM_RELU if relu else 0) | (_lib.GEMM_ACCUM if accum else 0) | \
            (_lib.GEMM_RELUMASK if mask else 0)
    call("occnerf_sgemm", A, sAi, sAr, B, sBr, sBj, Cp, ldc, bias, mask, ldmask, Mi, Nj, Kr, flags, split_k, stream(),
         work=2.0 * Mi * Nj * Kr)


class MlpSimt:
    name = "fp32"

    def forward(self, XB, raw, W: MlpWeights, save: bool):
        """XB [M,132] with columns 64..131 filled; writes XB[:, :64] and raw[:, 0:4].  Returns the saved state."""
        M, dev = XB.shape[0], XB.device
        xb, rw = XB.data_ptr(), raw.data_ptr()
        acts = []

        def buf(i):
            if save or len(acts) < 2:
                acts.append(torch.empty(M, 256, device=dev, dtype=f32))
                return acts[-1]
            return acts[i % 2]

        a_ptr, a_ld, K = xb + 4 * X0_OFF, XB_LD, 68
        for l in range(4):
            H = buf(l)
            _gemm(a_ptr, a_ld, 1, W.pts_w[l].data_ptr(), 1, K, H.data_ptr(), 256, M, 256, K, bias=W.pts_b[l].data_ptr(), relu=True)
            a_ptr, a_ld, K = H.data_ptr(), 256, 256
        gw, gb = W.geo_w.data_ptr(), W.geo_b.data_ptr()
        _gemm(a_ptr, 256, 1, gw, 1, 256, rw + 12, 5, M, 1, 256, bias=gb)                       # sigma -> raw[:,3]
        _gemm(a_ptr, 256, 1, gw + 4 * 256, 1, 256, xb, XB_LD, M, 64, 256, bias=gb + 4)         # geo feats -> XB[:, :64]
        w0 = _pad_rgb0(W.rgb_w[0])
        a_ptr, a_ld, K = xb, XB_LD, XB_LD
        for l in range(4):
            H = buf(4 + l)
            wl = w0 if l == 0 else W.rgb_w[l]
            _gemm(a_ptr, a_ld, 1, wl.data_ptr(), 1, K, H.data_ptr(), 256, M, 256, K, bias=W.rgb_b[l].data_ptr(), relu=True)
            a_ptr, a_ld, K = H.data_ptr(), 256, 256
        _gemm(a_ptr, 256, 1, W.out_w.data_ptr(), 1, 256, rw, 5, M, 3, 256, bias=W.out_b.data_ptr())
        return {"acts": acts, "w0": w0} if save else None

    def backward(self, XB, g_raw, W: MlpWeights, saved):
        """g_raw [M,5] -> gXB [M,132] (columns 64..131 = d(agg,var,h), both trunks summed) and the 20 parameter grads."""
        M, dev = XB.shape[0], XB.device
        acts, w0 = saved["acts"], saved["w0"]
        Hs, Rs = acts[0:4], acts[4:8]
        sk = _splits(M)
        gr, xb = g_raw.data_ptr(), XB.data_ptr()
        G = [torch.empty(M, 256, device=dev, dtype=f32) for _ in range(2)]
        gXB = torch.empty(M, XB_LD, device=dev, dtype=f32)
        gx = gXB.data_ptr()
        grads = {}

        def wgrad(name_w, name_b, dY, ldy, ny, X, ldx, kx, row0=0, out_w=None, out_b=None):
            """dW[row0:row0+ny, :kx] = dY^T X ; db = colsum(dY)."""
            gw = out_w if out_w is not None else torch.zeros(ny, kx, device=dev, dtype=f32)
            gb = out_b if out_b is not None else torch.zeros(ny, device=dev, dtype=f32)
            _gemm(dY, 1, ldy, X, ldx, 1, gw.data_ptr() + 4 * row0 * kx, kx, ny, kx, M, split_k=sk)
            call("occnerf_colsum", dY, ldy, None, 0, M, ny, gb.data_ptr() + 4 * row0, stream())
            grads[name_w], grads[name_b] = gw, gb

        # ---- colour trunk
        wgrad("out_w", "out_b", gr, 5, 3, Rs[3].data_ptr(), 256, 256)
        _gemm(gr, 5, 1, W.out_w.data_ptr(), 256, 1, G[0].data_ptr(), 256, M, 256, 3, mask=Rs[3].data_ptr(), ldmask=256)
        cur = 0
        for l in (3, 2, 1):
            wgrad(f"rgb_w{l}", f"rgb_b{l}", G[cur].data_ptr(), 256, 256, Rs[l - 1].data_ptr(), 256, 256)
            _gemm(G[cur].data_ptr(), 256, 1, W.rgb_w[l].data_ptr(), 256, 1, G[1 - cur].data_ptr(), 256, M, 256, 256,
                  mask=Rs[l - 1].data_ptr(), ldmask=256)
            cur = 1 - cur
        wgrad("rgb_w0p", "rgb_b0", G[cur].data_ptr(), 256, 256, xb, XB_LD, XB_LD)
        _gemm(G[cur].data_ptr(), 256, 1, w0.data_ptr(), XB_LD, 1, gx, XB_LD, M, XB_LD, 256)     # gXB = G_R1 . W'
        g0 = grads.pop("rgb_w0p")
        grads["rgb_w0"] = torch.cat([g0[:, :99], g0[:, 100:]], 1).contiguous()
        # ---- geometry head: dYg = (g_sigma, gXB[:, :64])
        gw, gb = torch.zeros(65, 256, device=dev, dtype=f32), torch.zeros(65, device=dev, dtype=f32)
        wgrad("geo_w", "geo_b", gr + 12, 5, 1, Hs[3].data_ptr(), 256, 256, row0=0, out_w=gw, out_b=gb)
        wgrad("geo_w", "geo_b", gx, XB_LD, 64, Hs[3].data_ptr(), 256, 256, row0=1, out_w=gw, out_b=gb)
        nxt = 1 - cur
        _gemm(gx, XB_LD, 1, W.geo_w.data_ptr() + 4 * 256, 256, 1, G[nxt].data_ptr(), 256, M, 256, 64)
        _gemm(gr + 12, 5, 1, W.geo_w.data_ptr(), 256, 1, G[nxt].data_ptr(), 256, M, 256, 1, accum=True,
              mask=Hs[3].data_ptr(), ldmask=256)
        cur = nxt
        # ---- geometry trunk
        for l in (3, 2, 1):
            wgrad(f"pts_w{l}", f"pts_b{l}", G[cur].data_ptr(), 256, 256, Hs[l - 1].data_ptr(), 256, 256)
            _gemm(G[cur].data_ptr(), 256, 1, W.pts_w[l].data_ptr(), 256, 1, G[1 - cur].data_ptr(), 256, M, 256, 256,
                  mask=Hs[l - 1].data_ptr(), ldmask=256)
            cur = 1 - cur
        wgrad("pts_w0", "pts_b0", G[cur].data_ptr(), 256, 256, xb + 4 * X0_OFF, XB_LD, 68)
        _gemm(G[cur].data_ptr(), 256, 1, W.pts_w[0].data_ptr(), 68, 1, gx + 4 * X0_OFF, XB_LD, M, 68, 256, accum=True)
        return gXB, [grads[k] for k in MlpWeights.ORDER]


def nonrigid_offsets(xyz, cond, window, nr_w, nr_b, const_off=None, return_const=False):
    """xyz (m,3) -> xyz + MLP([cond69, hann_pe36])  (mlp_offset.py:45-62), fp32, forward only (its output feeds
    no_grad code only).  When the Hann window is fully closed (every training iteration before
    kick_in_iter, network.py:579-583, where the condition code is zero as well) every row of the MLP input is the same
    [cond69, 0], so the offset is one constant 3-vector evaluated on a single row -- decided on the host from the
    window alone, the condition tensor is never inspected (no device->host sync, CUDA-graph capturable)."""
    m, dev = xyz.shape[0], xyz.device
    if const_off is not None:
        return xyz + const_off
    w = [t.detach().contiguous() for t in nr_w]
    b = [t.detach().contiguous() for t in nr_b]
    zero_in = all(v == 0.0 for v in window)
    rows = 1 if zero_in else m
    pe = torch.zeros(1, 36, device=dev, dtype=f32) if zero_in else ops.hann_pe(xyz, window)
    if cond is None:
        cond = torch.zeros(1, 69, device=dev, dtype=f32)
    cond = cond.reshape(1, 69).contiguous().float()
    # layer 0: fold the (row-independent) condition code into the bias
    b0 = torch.empty(1, 128, device=dev, dtype=f32)
    _gemm(cond.data_ptr(), 69, 1, w[0].data_ptr(), 1, 105, b0.data_ptr(), 128, 1, 128, 69, bias=b[0].data_ptr())
    H = [torch.empty(rows, 128, device=dev, dtype=f32) for _ in range(2)]
    _gemm(pe.data_ptr(), 36, 1, w[0].data_ptr() + 4 * 69, 1, 105, H[0].data_ptr(), 128, rows, 128, 36, bias=b0.data_ptr(), relu=True)
    cur = 0
    for l in (1, 2, 3):
        _gemm(H[cur].data_ptr(), 128, 1, w[l].data_ptr(), 1, 128, H[1 - cur].data_ptr(), 128, rows, 128, 128, bias=b[l].data_ptr(), relu=True)
        cur = 1 - cur
    # layer 4 takes [h128, pe36]
    _gemm(H[cur].data_ptr(), 128, 1, w[4].data_ptr(), 1, 164, H[1 - cur].data_ptr(), 128, rows, 128, 128)
    _gemm(pe.data_ptr(), 36, 1, w[4].data_ptr() + 4 * 128, 1, 164, H[1 - cur].data_ptr(), 128, rows, 128, 36, bias=b[4].data_ptr(),
          relu=True, accum=True)
    cur = 1 - cur
    _gemm(H[cur].data_ptr(), 128, 1, w[5].data_ptr(), 1, 128, H[1 - cur].data_ptr(), 128, rows, 128, 128, bias=b[5].data_ptr(), relu=True)
    cur = 1 - cur
    if zero_in:
        off = torch.empty(1, 3, device=dev, dtype=f32)
        _gemm(H[cur].data_ptr(), 128, 1, w[6].data_ptr(), 1, 128, off.data_ptr(), 3, 1, 3, 128, bias=b[6].data_ptr())
        return off if return_const else xyz + off
    out = xyz.clone()
    _gemm(H[cur].data_ptr(), 128, 1, w[6].data_ptr(), 1, 128, out.data_ptr(), 3, m, 3, 128, bias=b[6].data_ptr(), accum=True)
    return out
