"""tcgen05/TMEM engine of the canonical MLP (csrc/mlp_tc.cu) behind the same interface as mlp.MlpSimt.

forward : one fused kernel per chunk (occnerf_mlp_forward_tc), weights re-packed per call as bf16 (hi[,lo]) UMMA
          operand images (occnerf_mlp_pack_weights; 0.5 M parameters, microseconds).
backward: for now the exact-fp32 GEMM path of mlp.MlpSimt on the activations the fused forward saved (fp32), so a
          training step is "tensor-core forward + fp32 backward" until the fused dgrad/wgrad kernels land.
"""
from __future__ import annotations

import torch

from occnerf_b200 import _lib, mlp as M
from occnerf_b200._lib import call, stream

f32 = torch.float32


class MlpTc:
    def __init__(self, n_pass: int = 3):
        assert n_pass in (1, 3)
        self.n_pass = n_pass
        self.name = f"tc{n_pass}"
        self._simt = M.MlpSimt()

    def pack(self, W: M.MlpWeights, device):
        nbytes = _lib.load().occnerf_mlp_packed_bytes(self.n_pass)
        packed = torch.empty(nbytes, device=device, dtype=torch.uint8)
        P = _lib.MlpParams()
        ws = W.pts_w + [W.geo_w] + W.rgb_w + [W.out_w]
        bs = W.pts_b + [W.geo_b] + W.rgb_b + [W.out_b]
        for i in range(10):
            P.w[i], P.b[i] = ws[i].data_ptr(), bs[i].data_ptr()
        import ctypes
        call("occnerf_mlp_pack_weights", ctypes.byref(P), self.n_pass, packed.data_ptr(), stream())
        return packed

    def forward(self, XB, raw, W: M.MlpWeights, save: bool):
        m, dev = XB.shape[0], XB.device
        packed = self.pack(W, dev)
        acts = torch.empty(8, m, 256, device=dev, dtype=f32) if save else None
        call("occnerf_mlp_forward_tc", XB.data_ptr(), m, packed.data_ptr(), self.n_pass, raw.data_ptr(), raw.shape[1],
             acts.data_ptr() if save else None, 1 if save else 0, stream(), work=M.FLOP_FWD * m)
        if not save:
            return None
        return {"acts": [acts[i] for i in range(8)], "w0": M._pad_rgb0(W.rgb_w[0]), "packed": packed}

    def backward(self, XB, g_raw, W: M.MlpWeights, saved):
        return self._simt.backward(XB, g_raw, W, saved)
