"""CPU: host-side logic of the sharded view render (occnerf_b200/render.py): per-rank frames painted from contiguous
ray ranges merge into exactly the single-rank frame, for any world size (incl. more ranks than rays)."""
import os

import numpy as np
import pytest

from occnerf_b200 import distributed as D
from occnerf_b200 import render
from oracle import image_oracle as IO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_merge_frames_equals_single_rank(world):
    g = np.load(os.path.join(GOLDEN, "image_unpack.npz"))
    H, W, mask, bg = int(g["H"]), int(g["W"]), g["ray_mask"], g["bgcolor"] / 255.
    n = int(mask.sum())
    pix = np.nonzero(mask)[0]
    frames, spans = [], []
    for r in range(world):
        b, e = D.shard_range(n, r, world)
        own = np.zeros_like(mask)
        own[pix[b:e]] = True
        frames.append(IO.unpack(W, H, own, bg, g["rgb"][b:e], g["alpha"][b:e]))   # what rank r paints
        spans.append((b, e))
    rgb8, alpha8 = render.merge_frames(frames, mask, spans)
    assert np.array_equal(rgb8, g["rgb_image"]) and np.array_equal(alpha8, g["alpha_image"][..., 0])


def test_merge_frames_with_more_ranks_than_rays():
    mask = np.zeros(16, bool)
    mask[[3, 9]] = True
    rgb, alpha = np.array([[0.2, 0.4, 0.6], [1.0, 0.0, 0.5]], np.float32), np.array([0.5, 1.0], np.float32)
    full = IO.unpack(4, 4, mask, [0, 0, 0], rgb, alpha)
    frames, spans = [], []
    for r in range(4):
        b, e = D.shard_range(2, r, 4)
        own = np.zeros_like(mask)
        own[np.nonzero(mask)[0][b:e]] = True
        frames.append(IO.unpack(4, 4, own, [0, 0, 0], rgb[b:e], alpha[b:e]))
        spans.append((b, e))
    got = render.merge_frames(frames, mask, spans)
    assert np.array_equal(got[0], full[0]) and np.array_equal(got[1], full[1])


# ----------------------------------------------------------------------------- world_size 2, gloo
def _oracle_backed_ops(monkeypatch_target):
    """Stand-ins for the two CUDA entry points used by render.render_view, backed by the numpy oracles, so that the HOST
    logic of the sharded view render (ranges, empty shards, painting, merge) can run under a real 2-process gloo group on
    CPU.  Test scaffolding only: the product has no such path."""
    import torch
    from oracle import rays_oracle as RO

    def generate_rays(H, W, K, R, T, bmin, bmax, device="cpu", want_pixel_index=False, **_):
        packed, mask, pix = RO.frame_rays(H, W, K, R, T, bmin, bmax)
        return torch.from_numpy(packed), torch.from_numpy(mask), packed.shape[0], torch.from_numpy(pix)

    def unpack_image(rgb, alpha, pix, H, W, bg, out=None, fill=True):
        own = np.zeros(H * W, bool)
        own[pix.numpy()] = True
        order = np.argsort(pix.numpy(), kind="stable")
        r8, a8 = IO.unpack(W, H, own, bg, rgb.numpy()[order], alpha.numpy()[order])
        return torch.from_numpy(r8), torch.from_numpy(a8), torch.zeros(1, dtype=torch.int32)

    monkeypatch_target.generate_rays = generate_rays
    monkeypatch_target.unpack_image = unpack_image


class _FakeNet:
    """Per-ray outputs that depend only on the ray itself (like the real path: rays are independent)."""

    def parameters(self):
        import torch
        yield torch.zeros(1)

    def __call__(self, rays, near, far, iter_val, **data):
        import torch
        o, d = rays
        rgb = torch.sigmoid(d * 3.0 + o * 0.1)
        return {"rgb": rgb, "alpha": torch.clamp((far - near).reshape(-1) / 3.0, 0, 1)}


def _view_worker(rank, world, port, tmp):
    import pickle
    import torch.distributed as dist
    from occnerf_b200 import ops, synthetic as S
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _oracle_backed_ops(ops)
        H, W = 40, 56
        K, R, T = S.lookat_camera(W, yaw=0.5)
        box = {"min_xyz": np.array([-0.95, -1.35, -0.45], np.float32), "max_xyz": np.array([0.95, 0.65, 0.45], np.float32)}
        bg = [0.1, 0.5, 0.9]
        part = render.render_view(_FakeNet(), H, W, K, R.astype(np.float64), T.astype(np.float64), box, {}, bg, rank=rank, world=world)
        mine = (part["rgb8"].numpy(), part["alpha8"].numpy(), part["rays"])
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)                         # host-side assembly, outside any timed path
        if rank == 0:
            full = render.render_view(_FakeNet(), H, W, K, R.astype(np.float64), T.astype(np.float64), box, {}, bg, rank=0, world=1)
            n = full["rays"][2]
            spans = [g[2][:2] for g in gathered]
            assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            rgb8, alpha8 = render.merge_frames([(g[0], g[1]) for g in gathered], full["ray_mask"].numpy(), spans)
            assert n > 100 and np.array_equal(rgb8, full["rgb8"].numpy()) and np.array_equal(alpha8, full["alpha8"].numpy())
            with open(os.path.join(tmp, "ok"), "wb") as f:
                pickle.dump(n, f)
    finally:
        dist.destroy_process_group()


def test_two_rank_view_render_merges_to_the_single_rank_frame(tmp_path):
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_view_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()
