"""Shared test plumbing: rebuild a golden case's inputs from seeds and load the reference's outputs."""
import os

import numpy as np
import torch

from oracle import make_golden

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    sub, w, fr, vol, t_rand, rk = make_golden.build_case(name)
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"render_{name}.npz")))
    # the seeded generator must reproduce exactly what the fixture was made from
    assert np.array_equal(g["rays_d"], fr.rays_d.numpy()) and np.array_equal(g["near"], fr.near.numpy())
    assert np.array_equal(g["motion_scale_Rs"], fr.motion_scale_Rs.numpy())
    # (sums of 15.5 M values: the reduction order depends on the host's core count, so compare to 1e-9 relative)
    assert np.isclose(float(g["emb_checksum"]), w.embeddings.double().sum().item(), rtol=1e-9, atol=1e-9)
    assert np.isclose(float(g["vol_checksum"]), vol.double().sum().item(), rtol=1e-9, atol=1e-9)
    if t_rand is not None:
        assert np.array_equal(g["t_rand"], t_rand.numpy())
    return sub, w, fr, vol, t_rand, rk, g


def scatter_dense(idx, val, numel):
    out = torch.zeros(numel, dtype=torch.float32)
    out[torch.from_numpy(idx)] = torch.from_numpy(val)
    return out


def normwise_close(a, b, rel):
    """max|a-b| <= rel * max|b| : the right yardstick for summed gradients, whose small entries are
    cancellation noise in any summation order."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.abs(b).max()
    return bool(np.abs(a - b).max() <= rel * scale + 1e-30)


def dev():
    return torch.device("cuda:0")


def report(name, **vals):
    """Appends a line to gpurun_out/parity.log so that a GPU run leaves its measured errors behind."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    with open(os.path.join(root, "gpurun_out", "parity.log"), "a") as f:
        f.write(name + " " + " ".join(f"{k}={v:.3e}" if isinstance(v, float) else f"{k}={v}" for k, v in vals.items()) + "\n")


def maxabs(a, b):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max()) if a.size else 0.0
