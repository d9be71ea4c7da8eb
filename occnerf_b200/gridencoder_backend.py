"""Drop-in for the reference's `_gridencoder` pybind module (gridencoder/src/bindings.cpp:5-9): the same three
function names, positional argument order and in-place output conventions (gridencoder.h:12-15), so that the
reference's own `gridencoder/grid.py` binds to it unmodified:

    import sys, occnerf_b200.gridencoder_backend as be
    sys.modules["_gridencoder"] = be          # grid.py:9 does `import _gridencoder as _backend`

Tensors must be CUDA, contiguous, fp32 (offsets int32) exactly as the reference's CHECK_* macros demand
(gridencoder.cu:449-465); anything else raises RuntimeError.  Kernels run on torch's current stream.
"""
from __future__ import annotations

import numpy as np

from occnerf_b200 import _lib, ops
from occnerf_b200._lib import ptr

f32 = ops.f32


def _check(gridtype, align_corners, interp):
    if gridtype != 0 or align_corners or interp != 0:
        raise RuntimeError("occnerf_b200 _gridencoder: only gridtype=hash(0), align_corners=False, interp=linear(0) "
                           "(the configuration OccNeRF instantiates, occnerf_mlp.py:45) is built")


def grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype, align_corners, interp=0):
    """inputs [B,D] -> outputs [L,B,C] (pre-allocated, written in place); dy_dx [B, L*D*C] or None."""
    _check(gridtype, align_corners, interp)
    scales = ops.level_scales(float(S), int(H), int(L), inputs.device)
    _lib.call("occnerf_hashgrid_forward", ptr(inputs, f32), ptr(embeddings, f32), ptr(offsets, ops.i32), ptr(scales, f32),
              ptr(outputs, f32), _lib.LAYOUT_LBC, int(L * C), int(B), int(D), int(C), int(L), ptr(dy_dx, f32), None, None, 0,
              _lib.stream())


def grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx, grad_inputs, gridtype,
                         align_corners, interp=0):
    """grad [L,B,C]; grad_embeddings accumulated in place (caller zeroes, grid.py:78); grad_inputs [B,D] iff dy_dx."""
    _check(gridtype, align_corners, interp)
    scales = ops.level_scales(float(S), int(H), int(L), inputs.device)
    _lib.call("occnerf_hashgrid_backward", ptr(grad, f32), _lib.LAYOUT_LBC, int(L * C), ptr(inputs, f32), ptr(offsets, ops.i32),
              ptr(scales, f32), ptr(grad_embeddings, f32), int(B), int(D), int(C), int(L), 0, _lib.stream())
    if dy_dx is not None:
        _lib.call("occnerf_hashgrid_input_backward", ptr(grad, f32), _lib.LAYOUT_LBC, int(L * C), ptr(dy_dx, f32),
                  ptr(grad_inputs, f32), int(B), int(D), int(C), int(L), _lib.stream())


def grad_total_variation(inputs, embeddings, grad, offsets, weight, B, D, C, L, S, H, gridtype, align_corners):
    raise RuntimeError("grad_total_variation is never called on OccNeRF's path (GridEncoder.grad_total_variation has no "
                       "caller, SURVEY.md section 2.2) and is not built")
