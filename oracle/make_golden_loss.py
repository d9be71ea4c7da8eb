"""Writes tests/golden/loss.npz with the reference's own `_unpack_imgs` and `img2mse` (core/train/trainers/occnerf/trainer.py:24,31-41,
compiled from the source text: the module itself imports the whole training stack) on a seeded case with partially covered patches.
Run in the build container only."""
import ast
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("OCCNERF_REFERENCE", "/root/reference")


def reference_fns():
    src = open(os.path.join(REF, "core", "train", "trainers", "occnerf", "trainer.py")).read()
    ns = {"torch": torch, "np": np}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == "_unpack_imgs":
            exec(compile(ast.Module([node], []), "trainer.py", "exec"), ns)
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "img2mse":
            exec(compile(ast.Module([node], []), "trainer.py", "exec"), ns)
    return ns["_unpack_imgs"], ns["img2mse"]


def main():
    unpack, img2mse = reference_fns()
    g = torch.Generator().manual_seed(0)
    N, P, S = 5, 16, 8
    masks = torch.rand(N, P, P, generator=g) > 0.15
    masks[1] = True
    counts = masks.reshape(N, -1).sum(1)
    div = [0] + torch.cumsum(counts, 0).tolist()
    n = div[-1]
    rgbs = torch.rand(n, 3, generator=g, dtype=torch.float32).requires_grad_(True)
    comp = (torch.rand(n, S, generator=g) * 10).requires_grad_(True)
    targets = torch.rand(N, P, P, 3, generator=g)
    bg = torch.tensor([0.2, 0.5, 0.9])
    imgs = unpack(rgbs, masks, bg, targets, div)
    w_mse, w_comp = 0.2, 1.0
    total = w_mse * img2mse(imgs, targets) + w_comp * torch.mean(comp)
    total.backward()
    path = os.path.join(ROOT, "tests", "golden", "loss.npz")
    np.savez_compressed(path, patch_masks=masks.numpy(), div=np.array(div), rgbs=rgbs.detach().numpy(), comp=comp.detach().numpy(),
                        targets=targets.numpy(), bgcolor=bg.numpy(), w_mse=w_mse, w_comp=w_comp, patch_imgs=imgs.detach().numpy(),
                        loss=total.detach().numpy(), g_rgbs=rgbs.grad.numpy(), g_comp=comp.grad.numpy())
    print(path, float(total), n)


if __name__ == "__main__":
    main()
