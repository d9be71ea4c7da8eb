"""GPU: occnerf_clip_adam_step (global grad-norm clip + Adam, csrc/optim.cu) against the library calls it replaces in the
reference's trainer: torch.nn.utils.clip_grad_norm_(params, 1.0) + torch.optim.Adam.step() (trainer.py:248-249)."""
import pytest
import torch

from occnerf_b200.optim import ClipAdam
from tests.helpers import dev, report

pytestmark = pytest.mark.gpu


def _make(seed, shapes):
    gen = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=gen) * 0.1 for s in shapes]


@pytest.mark.parametrize("gscale", [10.0, 1e-3])          # clipping active / inactive
def test_matches_torch_clip_and_adam(gscale):
    shapes = [(7755336 // 8, 2), (256, 256), (256,), (65, 256), (3,), (6890, 1), (1,), (70000,)] + [(128, 128)] * 70     # > one batch of 64 tensors
    d = dev()
    ours = [torch.nn.Parameter(t.to(d)) for t in _make(0, shapes)]
    ref = [torch.nn.Parameter(t.to(d)) for t in _make(0, shapes)]
    lrs = [5e-4 if i % 3 else 1e-4 for i in range(len(shapes))]
    opt = ClipAdam([{"params": [p], "lr": lr, "name": str(i)} for i, (p, lr) in enumerate(zip(ours, lrs))], max_norm=1.0)
    topt = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(ref, lrs)], betas=(0.9, 0.999))
    worst = 0.0
    for step in range(4):
        grads = _make(100 + step, shapes)
        for i, (p, q, g) in enumerate(zip(ours, ref, grads)):
            g = g.to(d) * gscale
            if i == 0:
                g[::3] = 0.0                   # untouched table rows: exact zeros
            if i == 5 and step == 0:
                p.grad, q.grad = None, None     # a tensor without gradient in this step is skipped by both
                continue
            p.grad, q.grad = g.clone(), g.clone()
        norm = torch.nn.utils.clip_grad_norm_(ref, 1.0)
        topt.step()
        opt.step()
        assert abs(float(opt.grad_norm()) - float(norm)) <= 1e-5 * float(norm)
        for p, q in zip(ours, ref):
            worst = max(worst, float((p - q).abs().max()))
    report(f"clip_adam[gscale={gscale}]", worst_param_diff=worst)
    assert worst < 2e-7
    # gradients are left untouched (the clip coefficient is applied on the fly)
    assert torch.equal(ours[1].grad, (grads[1].to(d) * gscale))


def test_state_dict_round_trip_with_torch_adam():
    d = dev()
    shapes = [(300, 7), (11,)]
    a = [torch.nn.Parameter(t.to(d)) for t in _make(1, shapes)]
    b = [torch.nn.Parameter(t.to(d)) for t in _make(1, shapes)]
    topt = torch.optim.Adam([{"params": [p], "lr": 1e-3} for p in a])
    for step in range(3):
        for p, g in zip(a, _make(50 + step, shapes)):
            p.grad = g.to(d) * 1e-3
        topt.step()
    ours = ClipAdam([{"params": [p], "lr": 1e-3} for p in b], max_norm=1.0)
    with torch.no_grad():
        for p, q in zip(b, a):
            p.copy_(q)
    ours.load_state_dict(topt.state_dict())
    gs = _make(99, shapes)
    for p, q, g in zip(a, b, gs):
        p.grad, q.grad = g.to(d) * 1e-3, g.to(d) * 1e-3
    torch.nn.utils.clip_grad_norm_(a, 1.0)
    topt.step()
    ours.step()
    for p, q in zip(a, b):
        assert float((p - q).abs().max()) < 1e-7
    sd = ours.state_dict()
    assert set(sd) == {"state", "param_groups"} and float(sd["state"][0]["step"]) == 4.0


def test_capturable_in_a_cuda_graph():
    d = dev()
    p = torch.nn.Parameter(torch.ones(5000, device=d))
    p.grad = torch.full((5000,), 0.5, device=d)
    opt = ClipAdam([p], lr=1e-2, max_norm=1.0)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        opt.step()                       # allocates the state outside the capture
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        opt.step()
    before = p.detach().clone()
    g.replay()
    torch.cuda.synchronize()
    assert float((p - before).abs().max()) > 0.0 and float(opt.state[p]["step"]) == 2.0      # one eager step + one replay (the capture itself does not execute)
