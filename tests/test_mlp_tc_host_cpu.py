"""Host logic of the tensor-core MLP engine's backward orchestration (occnerf_b200/mlp_tc.py) with the C calls stubbed out: the
weight-gradient launches of the chunks of one query share ONE (dW, dB) buffer, `defer=True` hands nothing back until finish_wgrad, and
the 20 gradients come back in MlpWeights.ORDER with the reference's nn.Linear shapes (canonical_mlps/occnerf_mlp.py:24-77)."""
import torch

from occnerf_b200 import mlp as M, mlp_tc


def _engine(monkeypatch, calls):
    def fake_call(name, *args, **kw):
        calls.append((name, args))
    monkeypatch.setattr(mlp_tc, "call", fake_call)
    monkeypatch.setattr(mlp_tc, "stream", lambda: 0)
    monkeypatch.setattr(mlp_tc, "WGRAD_OVERLAP", False)      # (the side stream needs a GPU; the bookkeeping is the same)
    e = mlp_tc.MlpTc(n_pass=2)
    monkeypatch.setattr(e, "_packed", lambda *a, **k: torch.zeros(1))
    return e


def _saved(m):
    stride = (m + 63) // 64 * 64
    return {"acts": torch.zeros(10, 32, stride, 8, dtype=torch.bfloat16), "mask": torch.zeros(8, 32, stride, dtype=torch.uint8)}


def test_deferred_weight_gradients_share_one_buffer(monkeypatch):
    calls = []
    e = _engine(monkeypatch, calls)
    shared = {}
    for m, last in ((100, False), (64, False), (37, True)):
        gXB, grads = e.backward(torch.zeros(m, M.XB_LD), torch.zeros(m, 5), None, _saved(m), shared=shared, last=last, defer=True)
        assert grads is None and tuple(gXB.shape) == (m, M.XB_LD)
    wgrads = [a for n, a in calls if n == "occnerf_mlp_wgrad_tc"]
    assert len(wgrads) == 3 and [n for n, _ in calls].count("occnerf_mlp_backward_tc") == 3
    assert len({a[4] for a in wgrads}) == 1 and len({a[5] for a in wgrads}) == 1, "one dW / dB buffer for all chunks"
    assert [a[2] for a in wgrads] == [100, 64, 37] and [a[3] for a in wgrads] == [128, 64, 64]       # rows, padded stride
    out = e.finish_wgrad(shared)
    assert not shared, f"finish_wgrad must release the per-call state, left: {list(shared)}"
    shapes = [tuple(t.shape) for t in out]
    want = {"pts_w0": (256, 68), "pts_b0": (256,), "geo_w": (65, 256), "geo_b": (65,), "rgb_w0": (256, 131), "out_w": (3, 256), "out_b": (3,)}
    for name, shape in want.items():
        assert shapes[M.MlpWeights.ORDER.index(name)] == shape, (name, shapes[M.MlpWeights.ORDER.index(name)])
    assert len(out) == 20


def test_immediate_form_returns_gradients_with_the_last_chunk(monkeypatch):
    calls = []
    e = _engine(monkeypatch, calls)
    shared = {}
    _, g0 = e.backward(torch.zeros(64, M.XB_LD), torch.zeros(64, 5), None, _saved(64), shared=shared, last=False)
    _, g1 = e.backward(torch.zeros(64, M.XB_LD), torch.zeros(64, 5), None, _saved(64), shared=shared, last=True)
    assert g0 is None and len(g1) == 20 and "dW" not in shared
    _, g2 = e.backward(torch.zeros(64, M.XB_LD), torch.zeros(64, 5), None, _saved(64))            # no shared state: self-contained
    assert len(g2) == 20
