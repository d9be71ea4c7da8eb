"""CPU: the `_gridencoder` drop-in (occnerf_b200/gridencoder_backend.py) exposes the reference operator ABI --
the three function names of src/bindings.cpp:5-9 with the parameter names and order of src/gridencoder.h:12-15
(tests/golden/gridencoder_signatures.json is parsed from that header by oracle/make_golden_keys.py) -- and keeps the
reference's error behaviour: non-CUDA tensors and unsupported configurations raise RuntimeError (CHECK_CUDA,
gridencoder.cu:449-452; std::runtime_error :381,398).  No compute without a GPU."""
import inspect
import json
import os

import pytest
import torch

from occnerf_b200 import gridencoder_backend as be

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gridencoder_signatures.json")


def test_names_and_positional_order_match_the_reference_header():
    want = json.load(open(GOLDEN))
    assert sorted(want) == ["grad_total_variation", "grid_encode_backward", "grid_encode_forward"]
    for name, params in want.items():
        got = list(inspect.signature(getattr(be, name)).parameters)
        assert got == params, (name, got, params)


def _args(B=4):
    x = torch.rand(B, 4)
    emb = torch.zeros(100, 2)
    offs = torch.zeros(17, dtype=torch.int32)
    out = torch.empty(16, B, 2)
    return x, emb, offs, out


def test_cpu_tensors_are_rejected():
    x, emb, offs, out = _args()
    with pytest.raises(RuntimeError):
        be.grid_encode_forward(x, emb, offs, out, 4, 4, 2, 16, 0.5, 16, None, 0, False, 0)
    with pytest.raises(RuntimeError):
        be.grid_encode_backward(out, x, emb, offs, torch.zeros_like(emb), 4, 4, 2, 16, 0.5, 16, None, None, 0, False, 0)


@pytest.mark.parametrize("gridtype,align,interp", [(1, False, 0), (0, True, 0), (0, False, 1)])
def test_unsupported_configurations_raise(gridtype, align, interp):
    x, emb, offs, out = _args()
    with pytest.raises(RuntimeError, match="only gridtype=hash"):
        be.grid_encode_forward(x, emb, offs, out, 4, 4, 2, 16, 0.5, 16, None, gridtype, align, interp)


def test_total_variation_is_not_built():
    x, emb, offs, _ = _args()
    with pytest.raises(RuntimeError, match="never called"):
        be.grad_total_variation(x, emb, torch.zeros_like(emb), offs, 1.0, 4, 4, 2, 16, 0.5, 16, 0, False)
