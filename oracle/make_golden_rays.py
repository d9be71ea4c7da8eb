"""ORACLE tooling: write tests/golden/rays_*.npz with the UNMODIFIED reference functions get_rays_from_KRT and
rays_intersect_3d_bbox (core/utils/camera_util.py), imported from /root/reference in the build container.

    python -m oracle.make_golden_rays

cv2 and trimesh (imported at the top of camera_util.py, used by other functions only) are absent here and are
replaced by empty modules.  Two cameras: `zju` = float32 K with float64 extrinsics (what train.py:425-448 ends up
with after apply_global_tfm_to_camera) and `f64` = everything float64; both 64 x 48 pixels so that a row/column
mix-up cannot pass; `f32` = everything float32 (tpose.py:66-84).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("OCCNERF_REFERENCE_ROOT", "/root/reference")


def _camera_util():
    for name in ("cv2", "trimesh"):
        sys.modules.setdefault(name, types.ModuleType(name))
    spec = importlib.util.spec_from_file_location("_ref_camera_util", os.path.join(REF, "core", "utils", "camera_util.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cameras():
    rng = np.random.default_rng(42)
    out = {}
    for name, kdt in (("zju", np.float32), ("f64", np.float64), ("f32", np.float32)):
        H, W = 48, 64
        K = np.array([[156.25 * 1.03, 0.0, W / 2 + 0.37], [0.0, 156.25 * 0.98, H / 2 - 0.21], [0.0, 0.0, 1.0]], kdt)
        ax = rng.normal(size=3)
        ax /= np.linalg.norm(ax)
        ang = 0.12 if name == "f32" else 0.35      # (the third random axis at 0.35 rad turns the box out of the frame)
        Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        R = np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx
        campos = np.array([0.4, -0.25, 6.0]) + rng.normal(size=3) * 0.1
        T = -R @ campos
        bmin = np.array([-0.75, -1.1, -0.35], np.float32)
        bmax = np.array([0.80, 0.55, 0.30], np.float32)
        if name == "f32":                      # tpose.py:66-84 builds K and E in float32 -> numpy keeps float32 throughout
            R, T = R.astype(np.float32), T.astype(np.float32)
        out[name] = dict(H=H, W=W, K=K, R=R, T=T, bbox_min=bmin, bbox_max=bmax)
    return out


def _unpack_to_image():
    """run.py cannot be imported (it pulls the config system and the whole model); its two image functions are taken as
    they are written -- the function source text is exec'd unmodified -- over the real core/utils/image_util.py."""
    import ast
    sys.modules.setdefault("termcolor", types.SimpleNamespace(colored=lambda s, *a, **k: s))
    spec = importlib.util.spec_from_file_location("_ref_image_util", os.path.join(REF, "core", "utils", "image_util.py"))
    iu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(iu)
    src = open(os.path.join(REF, "run.py")).read()
    ns = {"np": np, "to_8b_image": iu.to_8b_image, "to_8b3ch_image": iu.to_8b3ch_image}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in ("unpack_alpha_map", "unpack_to_image"):
            exec(compile(ast.Module([node], []), "run.py", "exec"), ns)
    return ns["unpack_to_image"]


def image_case():
    g = np.load(os.path.join(ROOT, "tests", "golden", "rays_zju.npz"))
    rng = np.random.default_rng(7)
    mask = g["ray_mask"]
    n = int(mask.sum())
    rgb = rng.uniform(-0.2, 1.2, (n, 3)).astype(np.float32)
    rgb[:16] = np.linspace(0.0, 1.0, 48, dtype=np.float32).reshape(16, 3)          # exact grid values incl. 0 and 1
    alpha = rng.uniform(-0.1, 1.1, n).astype(np.float32)
    return dict(H=int(g["H"]), W=int(g["W"]), ray_mask=mask, rgb=rgb, alpha=alpha, bgcolor=np.array([255.0, 128.0, 7.0]))


def main():
    c = image_case()
    rgb_img, alpha_img, _ = _unpack_to_image()(c["W"], c["H"], c["ray_mask"], c["bgcolor"] / 255., c["rgb"], c["alpha"])
    path = os.path.join(ROOT, "tests", "golden", "image_unpack.npz")
    np.savez_compressed(path, **c, rgb_image=rgb_img, alpha_image=alpha_img)
    print(path, rgb_img.shape, rgb_img.dtype, alpha_img.shape)
    cu = _camera_util()
    for name, c in cameras().items():
        o, d = cu.get_rays_from_KRT(c["H"], c["W"], c["K"], c["R"], c["T"])
        o, d = o.reshape(-1, 3), d.reshape(-1, 3)
        near, far, mask = cu.rays_intersect_3d_bbox({"min_xyz": c["bbox_min"], "max_xyz": c["bbox_max"]}, o, d)
        path = os.path.join(ROOT, "tests", "golden", f"rays_{name}.npz")
        np.savez_compressed(path, **c, rays_o=o[mask].astype(np.float32), rays_d=d[mask].astype(np.float32),
                            near=near.astype(np.float32), far=far.astype(np.float32), ray_mask=mask,
                            rays_d_f64=d[mask], near_f64=near, far_f64=far)
        print(path, "valid rays", int(mask.sum()), "of", mask.size)


if __name__ == "__main__":
    main()
