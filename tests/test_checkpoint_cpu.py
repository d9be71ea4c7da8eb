"""CPU: the drop-in claim "reference checkpoints load" (trainer.py:398-415 saves {'network': state_dict}).
tests/golden/state_dict_keys.json lists name -> shape, dtype of the UNMODIFIED reference Network's state_dict
(oracle/make_golden_keys.py); occnerf_b200.network.Network must expose exactly that set, and a strict load of a
checkpoint with those entries must succeed and land in the tensors the kernels read."""
import json
import os

import torch

from occnerf_b200 import synthetic as S
from occnerf_b200.network import RenderConfig

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "state_dict_keys.json")


def _net():
    sub = S.make_subject(seed=0)
    net = S.network_from_synthetic(sub, S.make_weights(sub.bound, seed=0), RenderConfig(), device="cpu")
    net.install_prologue()
    return net


def test_state_dict_matches_the_reference_checkpoint_layout():
    want = json.load(open(GOLDEN))
    got = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in _net().state_dict().items()}
    assert sorted(got) == sorted(want), (sorted(set(want) - set(got)), sorted(set(got) - set(want)))
    assert got == want
    assert any(k.startswith("cnl_mlp.module.") for k in want)          # the nn.DataParallel level of the reference's keys


def test_strict_load_of_a_reference_shaped_checkpoint():
    want = json.load(open(GOLDEN))
    gen = torch.Generator().manual_seed(0)
    ckpt = {}
    for k, (shape, dtype) in want.items():
        dt = getattr(torch, dtype)
        ckpt[k] = torch.rand(shape, generator=gen).to(dt) if dt.is_floating_point else torch.zeros(shape, dtype=dt)
    net = _net()
    res = net.load_state_dict(ckpt, strict=True)                        # run.py:34
    assert not res.missing_keys and not res.unexpected_keys
    m = net.cnl_mlp.module
    assert torch.equal(m.encoder.embeddings, ckpt["cnl_mlp.module.encoder.embeddings"])
    assert torch.equal(net.point_dist, ckpt["point_dist"]) and torch.equal(net.point_base, ckpt["point_base"])


def test_network_api_surface_matches_the_reference():
    """Network.forward / _render_rays / _query_mlp / _batchify_rays take the reference's parameters, same names,
    same order (tests/golden/network_signatures.json, read from the reference's network.py by oracle/make_golden_keys.py);
    anything added after them is optional, and a reference method that swallows **kwargs still does."""
    import inspect
    from occnerf_b200.network import Network
    want = json.load(open(os.path.join(os.path.dirname(GOLDEN), "network_signatures.json")))
    for name, ref in want.items():
        params = list(inspect.signature(getattr(Network, name)).parameters.values())
        names = [p.name for p in params if p.kind == p.POSITIONAL_OR_KEYWORD]
        assert names[:len(ref["args"])] == ref["args"], (name, names, ref["args"])
        extras = [p for p in params if p.kind == p.POSITIONAL_OR_KEYWORD][len(ref["args"]):]
        assert all(p.default is not inspect.Parameter.empty for p in extras), (name, [p.name for p in extras])
        assert (ref["kwargs"] is not None) == any(p.kind == p.VAR_KEYWORD for p in params), name
    net = _net()
    assert net.deploy_mlps_to_secondary_gpus() is net                    # run.py:37 chains it


def test_hashgrid_level_table_matches_the_reference():
    """The level table of the reference's GridEncoder (grid.py:102-135; values in tests/golden/hashgrid_levels.json, read from
    the reference module itself): offsets are integer work -> identical, and so is per_level_scale (it enters exp2f)."""
    want = json.load(open(os.path.join(os.path.dirname(GOLDEN), "hashgrid_levels.json")))
    sub = S.make_subject(seed=0)
    assert abs(sub.bound - want["bound"]) < 1e-7
    offs, pls = S.hashgrid_offsets(desired_resolution=2048 * sub.bound)
    assert offs.tolist() == want["offsets"] and pls == want["per_level_scale"]
    enc = _net().cnl_mlp.module.encoder
    assert enc.offsets.tolist() == want["offsets"] and enc.offsets.dtype == torch.int32
    assert tuple(enc.embeddings.shape) == (want["offsets"][-1], want["level_dim"])
    assert 2 * want["offsets"][-1] == want["n_params"]
    assert float(enc.per_level_scale) == want["per_level_scale"]
