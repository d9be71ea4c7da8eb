"""View rendering around the ray path: what run.py:69-127 (`_freeview`) does for one batch, with the ray set-up
(freeview.py:208-219) and the image assembly (run.py:39-66) on the device, and the rays of a view sharded over ranks
(contiguous ray ranges, no data-path collective; BASELINE configs[2]).

    frame = render_view(net, H, W, K, R, T, dst_bbox, data, bgcolor, rank, world)

`data` holds what the reference's dataset puts in the batch besides the rays (dst_Rs, dst_Ts, cnl_gtfms,
motion_weights_priors, dst_posevec, cnl_bbox_min_xyz, cnl_bbox_scale_xyz, bgcolor) and is passed to `Network.forward`
unchanged.  Every rank evaluates the (cheap) ray set-up for the whole frame, renders its own contiguous range of the
valid rays and paints it into its own copy of the frame; `merge_frames` is the host-side assembly outside the timed path.
There is no CPU path: the kernels raise without a CUDA device.
"""
from __future__ import annotations

import numpy as np
import torch

from occnerf_b200 import distributed as D
from occnerf_b200 import ops


def view_rays(H, W, K, R, T, dst_bbox, device, rank: int = 0, world: int = 1):
    """-> (this rank's rays [m,8], its pixel indices [m] int32, ray_mask [H*W] bool, n valid rays of the whole view)."""
    rays, ray_mask, n, pix = ops.generate_rays(H, W, K, R, T, dst_bbox["min_xyz"], dst_bbox["max_xyz"], device=device,
                                               want_pixel_index=True)
    b, e = D.shard_range(n, rank, world)
    return rays[b:e], pix[b:e], ray_mask, n


def render_view(net, H, W, K, R, T, dst_bbox, data: dict, bgcolor_01, rank: int = 0, world: int = 1, iter_val=1e7):
    """One view (run.py:84-121).  Returns {'rgb8' [H,W,3] uint8, 'alpha8' [H,W] uint8, 'ray_mask', 'rays': (begin, end, n)};
    pixels of other ranks' rays hold the background until `merge_frames`."""
    dev = next(net.parameters()).device
    rays, pix, ray_mask, n = view_rays(H, W, K, R, T, dst_bbox, dev, rank, world)
    b, e = D.shard_range(n, rank, world)
    if rays.shape[0] > 0:
        with torch.no_grad():
            out = net(rays=(rays[:, 0:3], rays[:, 3:6]), near=rays[:, 6:7], far=rays[:, 7:8], iter_val=iter_val, **data)
        rgb, alpha = out["rgb"].reshape(-1, 3).float().contiguous(), out["alpha"].reshape(-1).float().contiguous()
    else:                                                   # a view that misses the box, or more ranks than rays
        rgb, alpha = torch.empty(0, 3, device=dev), torch.empty(0, device=dev)
    rgb8, alpha8, bad = ops.unpack_image(rgb, alpha, pix.contiguous(), H, W, bgcolor_01)
    return {"rgb8": rgb8, "alpha8": alpha8, "ray_mask": ray_mask, "rays": (b, e, n), "bad_pixels": bad}


def merge_frames(frames, ray_mask: np.ndarray, spans):
    """Host-side assembly of the per-rank frames of one view: rank r owns the pixels of the valid rays [b_r, e_r).
    frames: list of (rgb8 [H,W,3], alpha8 [H,W]) numpy arrays in rank order; spans: list of (b_r, e_r)."""
    pix = np.nonzero(np.asarray(ray_mask).reshape(-1))[0]
    rgb8, alpha8 = frames[0][0].copy().reshape(-1, 3), frames[0][1].copy().reshape(-1)
    for (f_rgb, f_alpha), (b, e) in zip(frames[1:], spans[1:]):
        own = pix[b:e]
        rgb8[own] = f_rgb.reshape(-1, 3)[own]
        alpha8[own] = f_alpha.reshape(-1)[own]
    return rgb8.reshape(frames[0][0].shape), alpha8.reshape(frames[0][1].shape)
