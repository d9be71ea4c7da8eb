"""CPU: the per-cell candidate lists of the grid KNN (occnerf_b200.ops.build_knn_grid) are supersets of the exact k nearest
neighbours of every query inside the cell -- the property that makes occnerf_knn_grid exact (include/occnerf_b200.h).
Host logic only (torch on the CPU); the kernel itself is compared with brute force in tests/test_knn_gpu.py."""
import torch

from occnerf_b200 import ops


def _cells(q, grid):
    gmin = torch.tensor(list(grid["params"])[:3])
    inv_h = grid["params"][3]
    dims = list(grid["dims"])
    f = (q - gmin) * inv_h
    inside = ((f >= 0) & (f < torch.tensor(dims, dtype=torch.float32))).all(1)
    c = f.floor().long()
    return (c[:, 2] * dims[1] + c[:, 1]) * dims[0] + c[:, 0], inside


def test_lists_are_supersets_of_the_exact_neighbours():
    gen = torch.Generator().manual_seed(0)
    base = torch.rand(400, 3, generator=gen) * torch.tensor([0.6, 0.9, 0.3])        # an elongated cloud
    base[100:110] = base[50:60]                                                    # duplicated points (ties)
    fps = [torch.randperm(400, generator=gen)[:90], torch.randperm(400, generator=gen)[:31], torch.randperm(400, generator=gen)[:12]]
    grid = ops.build_knn_grid(base, fps, cell=0.04, pad=0.12)
    assert grid["cells"] == grid["dims"][0] * grid["dims"][1] * grid["dims"][2]
    assert int(grid["cell_tab"][:, :, 1].sum()) == grid["entries"] == grid["lists"].numel()
    q = torch.cat([torch.rand(3000, 3, generator=gen) * torch.tensor([0.8, 1.1, 0.5]) - 0.1, base[:200], base[:50] + 1e-6])
    cell, inside = _cells(q, grid)
    assert inside.float().mean() > 0.9
    levels = [base, base[fps[0]], base[fps[1]], base[fps[2]]]
    for lev, P in enumerate(levels):
        k = min(10, P.shape[0])
        d = ((q[:, None, :] - P[None]) ** 2).sum(-1)
        kth = d.topk(k, largest=False)[0][:, -1]
        # every point at least as close as the k-th neighbour (ties included) must be in the cell's list
        need = d <= kth[:, None]
        for i in torch.nonzero(inside)[:, 0].tolist()[::7]:
            off, cnt = grid["cell_tab"][cell[i], lev].tolist()
            have = set(grid["lists"][off:off + cnt].tolist())
            want = set(torch.nonzero(need[i])[:, 0].tolist())
            assert want <= have, (lev, i, sorted(want - have))
        # lists are nearest-first with respect to the cell centre and never empty (k <= n points always qualify)
        assert int(grid["cell_tab"][:, lev, 1].min()) >= k


def test_lists_shrink_with_the_cell_size():
    gen = torch.Generator().manual_seed(1)
    base = torch.rand(300, 3, generator=gen) * 0.5
    fps = [torch.arange(0, 300, 4), torch.arange(0, 300, 16), torch.arange(0, 300, 64)]
    coarse = ops.build_knn_grid(base, fps, cell=0.08, pad=0.1)
    fine = ops.build_knn_grid(base, fps, cell=0.04, pad=0.1)
    mean = lambda g: float(g["cell_tab"][:, 0, 1].float().mean())
    assert mean(fine) < mean(coarse)


def test_lane_mapping_fills_the_warp():
    """ops.knn_grid_lane_rays: rays x samples per warp is always 32, the sample axis never exceeds the ray length, powers of two only."""
    from occnerf_b200 import ops
    for want in (1, 2, 4, 8, 16, 32):
        for stride in (1, 2, 3, 7, 8, 16, 31, 32, 128, 1000):
            lr = ops.knn_grid_lane_rays(want, stride)
            assert lr in (1, 2, 4, 8, 16, 32) and lr >= want
            assert 32 // lr <= max(1, stride), (want, stride, lr)
            if 32 // want <= stride:
                assert lr == want
    assert ops.knn_grid_lane_rays(2, 1) == 32 and ops.knn_grid_lane_rays(2, 128) == 2
