"""tcgen05/TMEM engine of the canonical MLP (csrc/mlp_tc.cu) behind the same interface as mlp.MlpSimt.

forward : one fused kernel per chunk (occnerf_mlp_forward_tc); weights re-packed per call as bf16 (hi[,lo]) UMMA
          operand images (occnerf_mlp_pack_weights; 0.5 M parameters, microseconds); post-ReLU activations saved as bf16.
backward: data gradients by the fused transposed chain (occnerf_mlp_backward_tc), which also leaves the bf16
          pre-activation gradients G_l behind; the ten weight gradients dW_l = G_l^T X_l and the bias gradients come
          from the MN-major tcgen05 kernel occnerf_mlp_wgrad_tc (csrc/mlp_wgrad.cu).
"""
from __future__ import annotations

import ctypes
import os

import torch

from occnerf_b200 import _lib, mlp as M
from occnerf_b200._lib import call, stream

f32, bf16 = torch.float32, torch.bfloat16

# Weight gradients beside the rest of the chunk's backward pass: with `defer=True` MlpTc.backward launches occnerf_mlp_wgrad_tc on a side
# stream (forked from the current one after the data-gradient chain) and MlpTc.finish_wgrad joins it.  The kernel is HBM-bound and holds
# no resources the scatter kernels behind the data gradients need (hash-grid / aggregation backward: L2 atomics, no shared memory), so
# they run on the same SMs at the same time.  OCCNERF_WGRAD_OVERLAP=0 (or WGRAD_OVERLAP = False) keeps everything on one stream.
WGRAD_OVERLAP = os.environ.get("OCCNERF_WGRAD_OVERLAP", "1") != "0"
_SIDE_STREAMS = {}


def _side_stream(dev):
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=key)
    return _SIDE_STREAMS[key]


class MlpTc:
    def __init__(self, n_pass: int = 3, wgrad: str = "tc", bwd_pass: int | None = None, pair: bool | None = None):
        assert n_pass in (1, 2, 3) and wgrad in ("tc", "lib") and bwd_pass in (None, 1, 2, 3)
        self.n_pass = n_pass
        # precision of the data-gradient chain: by default that of the forward chain; 1 = bf16 operands (what the
        # weight-gradient kernel uses anyway), independent of a split-bf16 forward
        self.bwd_pass = n_pass if bwd_pass is None else bwd_pass
        self.wgrad = wgrad      # "tc": hand-written tcgen05 kernel; "lib": cuBLAS (test cross-check only)
        # cta_group::2 CTA pairs (each CTA streams / holds half of the weight rows) or cta_group::1 CTAs sharing the stream by multicast
        self.pair = int(os.environ.get("OCCNERF_MLP_PAIR", "1")) != 0 if pair is None else bool(pair)
        self.name = "tf32" if n_pass == 2 else f"tc{n_pass}"      # 1 bf16 | 2 tf32 (kind::tf32) | 3 split-bf16

    def pack(self, W: M.MlpWeights, device, chain: int):
        nbytes = _lib.load().occnerf_mlp_packed_bytes(self.n_pass, chain)
        packed = torch.empty(nbytes, device=device, dtype=torch.uint8)
        P = _lib.MlpParams()
        ws = W.pts_w + [W.geo_w] + W.rgb_w + [W.out_w]
        bs = W.pts_b + [W.geo_b] + W.rgb_b + [W.out_b]
        for i in range(10):
            P.w[i], P.b[i] = ws[i].data_ptr(), bs[i].data_ptr()
        call("occnerf_mlp_pack_weights", ctypes.byref(P), self.n_pass, chain, int(self.pair), packed.data_ptr(), stream())
        return packed

    def _packed(self, W, dev, chain, n_pass, shared):
        """Packed operand images of the current weights; built once per _query_mlp call (`shared`) instead of once per chunk."""
        key = ("packed", chain, n_pass, self.pair)
        if shared is not None and key in shared:
            return shared[key]
        keep, self.n_pass = self.n_pass, n_pass
        packed = self.pack(W, dev, chain)
        self.n_pass = keep
        if shared is not None:
            shared[key] = packed
        return packed

    def forward(self, XB, raw, W: M.MlpWeights, save: bool, shared=None):
        m, dev = XB.shape[0], XB.device
        packed = self._packed(W, dev, 0, self.n_pass, shared)
        stride = (m + 63) // 64 * 64
        acts, mask = None, None
        if save:
            # chunk-major [slot][col/8][row][8] (include/occnerf_b200.h): coalesced epilogue stores, TMA-ready for wgrad
            acts = torch.empty(10, 32, stride, 8, device=dev, dtype=bf16)
            mask = torch.empty(8, 32, stride, device=dev, dtype=torch.uint8)      # one bit per hidden unit: ReLU'(x)
            if stride > m:
                acts[:, :, m:].zero_()
        call("occnerf_mlp_forward_tc", XB.data_ptr(), m, packed.data_ptr(), self.n_pass, int(self.pair), raw.data_ptr(), raw.shape[1],
             acts.data_ptr() if save else None, 2 if save else 0, stride, mask.data_ptr() if save else None, stream(),
             work=M.FLOP_FWD * m)
        return {"acts": acts, "mask": mask} if save else None

    def backward(self, XB, g_raw, W: M.MlpWeights, saved, shared=None, last=True, defer=False):
        """-> (gXB, list of the 20 parameter gradients | None).  With `shared` (a dict living as long as one _query_mlp
        call) the weight gradients of all its chunks accumulate in ONE dW/dB buffer -- the kernel adds into it anyway --
        and only the chunk with last=True maps it back to the nn.Linear layout; the others return None (= zero).
        defer=True (needs `shared`): always returns (gXB, None); the weight-gradient kernel may still be running on the side stream
        and the caller fetches the 20 gradients with finish_wgrad(shared) after the last chunk."""
        m, dev = XB.shape[0], XB.device
        acts = saved["acts"]
        stride = acts.shape[2]
        packed = self._packed(W, dev, 1, self.bwd_pass, shared)
        gXB = torch.empty(m, M.XB_LD, device=dev, dtype=f32)
        g_save = torch.empty(10, 32, stride, 8, device=dev, dtype=bf16)
        if stride > m:
            g_save[:, :, m:].zero_()
        call("occnerf_mlp_backward_tc", g_raw.data_ptr(), m, packed.data_ptr(), self.bwd_pass, int(self.pair), saved["mask"].data_ptr(), gXB.data_ptr(),
             g_save.data_ptr(), stride, stream(), work=M.FLOP_FWD * m)
        if self.wgrad == "lib":
            assert not defer, "the library weight-gradient cross-check has no deferred form"
            with _lib.region("lib:wgrad(cuBLAS bf16)+bias sums"):
                def rows(t):     # chunk-major -> row-major [slot][row][256]
                    return t.permute(0, 2, 1, 3).reshape(10, stride, 256)[:, :m]
                return gXB, self._wgrad_lib(XB, g_raw, rows(acts), rows(g_save))
        if shared is not None and "dW" in shared:
            dW, dB = shared["dW"], shared["dB"]
        else:
            dW = torch.zeros(10, 256, 256, device=dev, dtype=f32)
            dB = torch.zeros(10, 256, device=dev, dtype=f32)
            if shared is not None:
                shared["dW"], shared["dB"] = dW, dB
        if defer:
            assert shared is not None, "defer=True needs the per-call `shared` dict"
            if WGRAD_OVERLAP:
                side, cur = _side_stream(dev), torch.cuda.current_stream()
                keep = shared.setdefault("wgrad_keep", [])
                if len(keep) >= 3:          # long queries (many chunks): release the operands of finished launches every few chunks
                    cur.wait_stream(side)   # (their kernels ended during the data-gradient chain that has just been queued)
                    keep.clear()
                side.wait_stream(cur)                                          # fork: the chain's outputs are complete for the side stream
                with torch.cuda.stream(side):
                    call("occnerf_mlp_wgrad_tc", g_save.data_ptr(), acts.data_ptr(), m, stride, dW.data_ptr(), dB.data_ptr(), stream(),
                         work=M.FLOP_FWD * m)
                # the operands stay allocated until the join: the caching allocator would otherwise hand their memory to later work
                # of the main stream while the side stream still reads it
                keep.append((g_save, acts))
                shared["wgrad_side"] = side
            else:
                call("occnerf_mlp_wgrad_tc", g_save.data_ptr(), acts.data_ptr(), m, stride, dW.data_ptr(), dB.data_ptr(), stream(),
                     work=M.FLOP_FWD * m)
            return gXB, None
        call("occnerf_mlp_wgrad_tc", g_save.data_ptr(), acts.data_ptr(), m, stride, dW.data_ptr(), dB.data_ptr(), stream(),
             work=M.FLOP_FWD * m)
        if shared is not None and not last:
            return gXB, None
        if shared is not None:
            shared.pop("dW"), shared.pop("dB")
        return gXB, self._unpack_grads(dW, dB)

    def finish_wgrad(self, shared):
        """Join the side stream of the deferred weight-gradient launches of one _query_mlp call -> the 20 parameter gradients."""
        side = shared.pop("wgrad_side", None)
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        shared.pop("wgrad_keep", None)
        return self._unpack_grads(shared.pop("dW"), shared.pop("dB"))

    @staticmethod
    def _unpack_grads(dW, dB):
        g = {}
        g["pts_w0"], g["pts_b0"] = dW[0][:, :68].contiguous(), dB[0]
        for l in (1, 2, 3):
            g[f"pts_w{l}"], g[f"pts_b{l}"] = dW[l], dB[l]
        g["geo_w"] = torch.cat([dW[4][64:65], dW[4][:64]], 0).contiguous()          # sigma row back to the front
        g["geo_b"] = torch.cat([dB[4][64:65], dB[4][:64]], 0)
        g["rgb_w0"] = torch.cat([dW[5][:, :99], dW[5][:, 100:132]], 1).contiguous()  # drop the variance column
        g["rgb_b0"] = dB[5]
        for l in (1, 2, 3):
            g[f"rgb_w{l}"], g[f"rgb_b{l}"] = dW[5 + l], dB[5 + l]
        g["out_w"], g["out_b"] = dW[9][:3].contiguous(), dB[9][:3].contiguous()
        return [g[k] for k in M.MlpWeights.ORDER]

    def _wgrad_lib(self, XB, g_raw, acts, g_save):
        """Library fallback-free alternative for cross-checking the hand-written kernel: the same GEMMs through cuBLAS."""
        Xb = XB.to(bf16)
        g_out = g_raw[:, :3].to(bf16)

        def dw(G, X):
            return torch.matmul(G.t(), X).float()

        def db(G):
            return G.sum(0, dtype=f32)

        grads = {}
        grads["out_w"], grads["out_b"] = dw(g_out, acts[7]), g_raw[:, :3].sum(0)
        for i, l in enumerate((3, 2, 1)):
            grads[f"rgb_w{l}"], grads[f"rgb_b{l}"] = dw(g_save[i], acts[3 + l]), db(g_save[i])
        g0 = dw(g_save[3], Xb)
        grads["rgb_w0"], grads["rgb_b0"] = torch.cat([g0[:, :99], g0[:, 100:]], 1).contiguous(), db(g_save[3])
        Gg = g_save[4][:, :65]
        gw, gb = dw(Gg, acts[3]), db(Gg)
        grads["geo_w"], grads["geo_b"] = torch.cat([gw[64:65], gw[:64]], 0).contiguous(), torch.cat([gb[64:65], gb[:64]], 0)
        for i, l in enumerate((3, 2, 1)):
            grads[f"pts_w{l}"], grads[f"pts_b{l}"] = dw(g_save[5 + i], acts[l - 1]), db(g_save[5 + i])
        grads["pts_w0"], grads["pts_b0"] = dw(g_save[8], Xb[:, 64:132]), db(g_save[8])
        return [grads[k] for k in M.MlpWeights.ORDER]
