"""ORACLE (test infrastructure): ctypes front-end of oracle/hashgrid_oracle.c.

Exposes the C restatement of gridencoder.cu (a) as plain functions on CPU tensors and (b) as a
module object with the reference's pybind surface (`grid_encode_forward`, `grid_encode_backward`,
`grad_total_variation`; core/nets/occnerf/gridencoder/src/bindings.cpp:5-9) so that the reference's own
grid.py can run unmodified on CPU inside oracle/ref_shim.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle_hashgrid.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u32p = ctypes.POINTER(ctypes.c_uint32)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "hashgrid_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.hg_forward.argtypes = [_f32p, _f32p, _i32p, _f32p] + [ctypes.c_uint32] * 4 + [ctypes.c_float, ctypes.c_uint32,
                                    _f32p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, _f32p, ctypes.c_int, _u32p, _u32p]
        _lib.hg_backward.argtypes = [_f32p, _f32p, _i32p, _f32p, _f64p] + [ctypes.c_uint32] * 4 + [ctypes.c_float, ctypes.c_uint32,
                                     ctypes.c_uint32, ctypes.c_int, ctypes.c_int, _f32p, ctypes.c_int]
        _lib.hg_input_backward.argtypes = [_f32p, _f32p, _f32p] + [ctypes.c_uint32] * 4 + [ctypes.c_int]
        _lib.hg_host_level_scales.argtypes = [ctypes.c_float, ctypes.c_uint32, ctypes.c_uint32, _f32p]
        _lib.hg_set_threads.argtypes = [ctypes.c_int]
    return _lib


def set_threads(n: int) -> None:
    lib().hg_set_threads(int(n))


def _p(t, ty):
    if t is None:
        return ctypes.cast(None, ty)
    assert t.is_contiguous() and t.device.type == "cpu"
    return ctypes.cast(t.data_ptr(), ty)


def host_level_scales(S: float, H: int, L: int) -> torch.Tensor:
    out = torch.empty(L, dtype=torch.float32)
    lib().hg_host_level_scales(float(np.float32(S)), H, L, _p(out, _f32p))
    return out


def forward(inputs, emb, offsets, S, H, *, want_dy_dx=False, level_scales=None, lbc=False,
            want_cells=False, gridtype=0, align_corners=False, interp=0):
    """Returns dict(out, dy_dx, cells, idx).  out is [L,B,C] if lbc else [B,L*C]."""
    inputs = inputs.contiguous().float()
    B, D = inputs.shape
    C = emb.shape[1]
    L = offsets.shape[0] - 1
    out = torch.empty((L, B, C) if lbc else (B, L * C), dtype=torch.float32)
    dy_dx = torch.empty(B, L * D * C, dtype=torch.float32) if want_dy_dx else None
    cells = torch.empty(B, L, D, dtype=torch.int32) if want_cells else None
    idx = torch.empty(B, L, 1 << D, dtype=torch.int32) if want_cells else None
    lib().hg_forward(_p(inputs, _f32p), _p(emb, _f32p), _p(offsets, _i32p), _p(out, _f32p), B, D, C, L,
                     float(np.float32(S)), H, _p(dy_dx, _f32p), gridtype, int(align_corners), interp,
                     _p(level_scales, _f32p), int(lbc), _p(cells, _u32p), _p(idx, _u32p))
    return {"out": out, "dy_dx": dy_dx, "cells": cells, "idx": idx}


def backward(grad, inputs, offsets, n_entries, C, S, H, *, level_scales=None, lbc=False, want_f64=False,
             gridtype=0, align_corners=False, interp=0, grad_emb=None):
    inputs = inputs.contiguous().float()
    grad = grad.contiguous().float()
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    if grad_emb is None:
        grad_emb = torch.zeros(n_entries, C, dtype=torch.float32)
    g64 = torch.zeros(n_entries, C, dtype=torch.float64) if want_f64 else None
    lib().hg_backward(_p(grad, _f32p), _p(inputs, _f32p), _p(offsets, _i32p), _p(grad_emb, _f32p), _p(g64, _f64p),
                      B, D, C, L, float(np.float32(S)), H, gridtype, int(align_corners), interp,
                      _p(level_scales, _f32p), int(lbc))
    return grad_emb, g64


def input_backward(grad, dy_dx, B, D, C, L, lbc=False):
    gi = torch.empty(B, D, dtype=torch.float32)
    grad, dy_dx = grad.contiguous().float(), dy_dx.contiguous().float()   # keep the temporaries alive across the call
    lib().hg_input_backward(_p(grad, _f32p), _p(dy_dx, _f32p), _p(gi, _f32p), B, D, C, L, int(lbc))
    return gi


class HashGridFn(torch.autograd.Function):
    """Differentiable CPU hash-grid encode: [B,D] in [0,1] -> [B, L*C]  (grid.py:24-90 semantics)."""

    @staticmethod
    def forward(ctx, inputs, emb, offsets, S, H, level_scales):
        need_in = inputs.requires_grad
        r = forward(inputs.detach(), emb.detach(), offsets, S, H, want_dy_dx=need_in, level_scales=level_scales)
        ctx.save_for_backward(inputs.detach(), offsets, r["dy_dx"] if need_in else torch.empty(0))
        ctx.meta = (emb.shape[0], emb.shape[1], S, H, level_scales, need_in)
        return r["out"]

    @staticmethod
    def backward(ctx, g):
        inputs, offsets, dy_dx = ctx.saved_tensors
        n, C, S, H, ls, need_in = ctx.meta
        ge, _ = backward(g, inputs, offsets, n, C, S, H, level_scales=ls)
        gi = None
        if need_in:
            B, D = inputs.shape
            gi = input_backward(g, dy_dx, B, D, C, offsets.shape[0] - 1)
        return gi, ge, None, None, None, None


class RefBackendModule:
    """Drop-in for the reference's `_gridencoder` pybind module, CPU tensors (bindings.cpp:5-9)."""

    @staticmethod
    def grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype, align_corners, interp):
        r = forward(inputs, embeddings, offsets, S, H, want_dy_dx=dy_dx is not None, lbc=True,
                    gridtype=gridtype, align_corners=align_corners, interp=interp)
        outputs.copy_(r["out"])
        if dy_dx is not None:
            dy_dx.copy_(r["dy_dx"])

    @staticmethod
    def grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx, grad_inputs,
                             gridtype, align_corners, interp):
        backward(grad, inputs, offsets, embeddings.shape[0], C, S, H, lbc=True, gridtype=gridtype,
                 align_corners=align_corners, interp=interp, grad_emb=grad_embeddings)
        if dy_dx is not None:
            grad_inputs.copy_(input_backward(grad, dy_dx, B, D, C, L, lbc=True))

    @staticmethod
    def grad_total_variation(*_a, **_k):
        raise NotImplementedError("never called on the hot path (SURVEY.md section 2.2)")
