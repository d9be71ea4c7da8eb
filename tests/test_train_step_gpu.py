"""GPU: the CUDA-graph training step (occnerf_b200/train_step.py) replays exactly what the eager step does."""
import copy

import pytest
import torch

from occnerf_b200 import synthetic as S
from occnerf_b200.network import RenderConfig
from occnerf_b200.train_step import GraphedTrainStep
from tests.helpers import dev

pytestmark = pytest.mark.gpu


def _setup(seed=0):
    d = dev()
    sub = S.make_subject(seed=0)
    w = S.make_weights(sub.bound, seed=0)
    net = S.network_from_synthetic(sub, w, RenderConfig(perturb=0.0, mlp_engine="fp32"), device=d)   # no jitter: deterministic
    net.train(True)
    net.install_prologue()
    fr = S.make_frame(sub, mode="patch", n_patches=2, patch=16, seed=5)
    target = torch.rand(fr.rays_o.shape[0], 3, generator=torch.Generator().manual_seed(1))
    host = {k: v.pin_memory() for k, v in dict(rays_o=fr.rays_o, rays_d=fr.rays_d, near=fr.near, far=fr.far, dst_Rs=fr.dst_Rs,
            dst_Ts=fr.dst_Ts, cnl_gtfms=fr.cnl_gtfms, priors=sub.priors, posevec=fr.dst_posevec, bmin=fr.cnl_bbox_min_xyz,
            bscale=fr.cnl_bbox_scale_xyz, bg=fr.bgcolor, target=target).items()}
    return net, host


def _loss(out, d):
    return 0.2 * torch.mean((out["rgb"] - d["target"]) ** 2) + out["comp_loss"].mean()


def test_graph_replay_matches_eager_steps():
    d = dev()
    net_e, host = _setup()
    net_g = copy.deepcopy(net_e)
    net_g._cache = None
    assert net_g.mweight_vol_decoder is not net_e.mweight_vol_decoder
    steps = 3
    # eager reference: the same iteration body, run step by step
    params_e = [p for p in net_e.parameters() if p.requires_grad]
    opt_e = torch.optim.Adam(params_e, lr=5e-4, fused=True, capturable=True)
    eager = GraphedTrainStep.__new__(GraphedTrainStep)
    eager.net, eager.opt, eager.loss_fn, eager.iter_val, eager.max_norm, eager.params = net_e, opt_e, _loss, 500, 1.0, params_e
    eager.static = {k: v.to(d) for k, v in host.items()}
    eager.loss_dev, eager.grad_sync, eager._hits = torch.zeros(1, device=d), None, None
    # the graphed step warms up with 3 real iterations before capture: give the eager model the same head start
    losses_e = []
    for i in range(3 + steps):
        eager._iteration()
        losses_e.append(float(eager.loss_dev))
    params_g = [p for p in net_g.parameters() if p.requires_grad]
    opt_g = torch.optim.Adam(params_g, lr=5e-4, fused=True, capturable=True)
    gs = GraphedTrainStep(net_g, opt_g, _loss, host, 500, params=params_g, warmup=3)
    assert gs.launches > 20, "the capture must contain the library's kernels"
    losses_g = []
    for i in range(steps):
        lh = gs.step(host)
        torch.cuda.synchronize()
        losses_g.append(float(lh))
    # float atomics make two runs of the same step differ in the last bits and Adam (lr-sized steps whatever the gradient's
    # magnitude) amplifies that from step to step: observed 1e-4..3e-4 after the three warm-up steps, hence 5e-3
    for a, b in zip(losses_e[3:], losses_g):
        assert abs(a - b) <= 5e-3 * max(abs(a), 1e-6), (losses_e, losses_g)
    assert losses_g[-1] != losses_g[0], "replays must advance the optimisation"
    for (n, pe), pg in zip(net_e.named_parameters(), net_g.parameters()):
        if n == "point_counter":       # integer votes: a ray whose depth sits at the 0.5 threshold may vote in one run only
            assert float((pe != pg).float().mean()) < 0.01, n
            continue
        assert torch.allclose(pe, pg, rtol=1e-2, atol=2 * 5e-4 * (3 + steps)), n     # within a couple of Adam steps (lr = 5e-4)
    assert not gs.needs_recapture(501) and gs.needs_recapture(net_g.cfg.non_rigid_kick_in_iter)


def test_two_graph_step_with_sync_hook_and_native_optimizer():
    """Data-parallel shape of the step on one GPU: forward+backward graph, an eager hook on the static gradient tensors (what the
    NCCL all-reduce does between the graphs), optimizer graph with the native clip + Adam (occnerf_clip_adam_step)."""
    from occnerf_b200.optim import ClipAdam
    net_a, host = _setup()
    net_b = copy.deepcopy(net_a)
    net_b._cache = None
    calls = []

    def hook(grads, hits):                        # a "collective" that leaves the values alone (world size 1)
        calls.append(len([g for g in grads if g is not None]))
        for g in grads:
            if g is not None:
                g.mul_(1.0)

    pa = [p for p in net_a.parameters() if p.requires_grad]
    pb = [p for p in net_b.parameters() if p.requires_grad]
    one = GraphedTrainStep(net_a, ClipAdam(pa, lr=5e-4, max_norm=1.0), _loss, host, 500, params=pa, max_norm=None)
    two = GraphedTrainStep(net_b, ClipAdam(pb, lr=5e-4, max_norm=1.0), _loss, host, 500, params=pb, max_norm=None, grad_sync=hook)
    assert two.graph_opt is not None and one.graph_opt is None
    la, lb = [], []
    for _ in range(3):
        a, b = one.step(host), two.step(host)
        torch.cuda.synchronize()
        la.append(float(a)); lb.append(float(b))
    assert len(calls) >= 3 + 3 and calls[-1] > 20
    for a, b in zip(la, lb):
        assert abs(a - b) <= 5e-3 * max(abs(a), 1e-6), (la, lb)
    assert lb[-1] != lb[0]


def test_capture_after_non_rigid_kick_in():
    """with iter_val >= non_rigid_kick_in_iter `forward` hands the pose vector to `_query_mlp`; nothing on that path may
    synchronise with the host, or the step cannot be captured."""
    net, host = _setup()
    it = net.cfg.non_rigid_kick_in_iter + 1000
    params = [p for p in net.parameters() if p.requires_grad]
    from occnerf_b200.optim import ClipAdam
    gs = GraphedTrainStep(net, ClipAdam(params, lr=5e-4), _loss, host, it, params=params, max_norm=None, warmup=2)
    lh = gs.step(host); torch.cuda.synchronize(); l0 = float(lh)
    lh = gs.step(host); torch.cuda.synchronize(); l1 = float(lh)
    assert l0 == l0 and l1 == l1 and gs.needs_recapture(it + 1) and not gs.needs_recapture(it)
