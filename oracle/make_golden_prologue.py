"""ORACLE tooling: write tests/golden/prologue.npz -- the per-frame prologue of Network.forward (network.py:551-597: pose
refinement, MotionBasisComputer, MotionWeightVolumeDecoder) evaluated by the UNMODIFIED reference modules on CPU, with
the parameters of occnerf_b200.prologue.Prologue (seeded, see `seeded_prologue`) loaded into them by name.

    python -m oracle.make_golden_prologue

The 63.6 M decoder parameters are not stored: the test re-creates them from the seed, so the fixture holds only inputs
digests and the reference's outputs (the 25 x 32^3 volume sub-sampled 4x per axis plus per-channel sums)."""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occnerf_b200 import synthetic as S  # noqa: E402
from occnerf_b200 import prologue as P  # noqa: E402

PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "prologue.npz")


def seeded_prologue(seed: int = 123) -> P.Prologue:
    """Our prologue with reproducible, non-degenerate parameters (the stock init leaves the pose refiner's last layer
    almost at zero, which would make the refinement branch a no-op)."""
    torch.manual_seed(seed)
    pro = P.Prologue(pose_kick_in_iter=1000)
    gen = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for p in pro.pose_decoder.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=gen))
    return pro.eval()


def inputs():
    sub = S.make_subject(seed=0)
    fr = S.make_frame(sub, mode="patch", n_patches=1, patch=8, seed=21)
    return sub, fr


def main():
    warnings.filterwarnings("ignore", category=FutureWarning)
    from oracle import ref_shim
    sub, fr = inputs()
    pro = seeded_prologue()
    ref = ref_shim.build_reference_network(sub, S.make_weights(sub.bound, seed=0))
    for name in ("motion_basis_computer", "mweight_vol_decoder", "pose_decoder"):
        res = getattr(ref, name).load_state_dict(getattr(pro, name).state_dict(), strict=True)
        assert not res.missing_keys and not res.unexpected_keys
    out = {}
    with torch.no_grad():
        dst_Rs, dst_Ts, gt = fr.dst_Rs[None], fr.dst_Ts[None], fr.cnl_gtfms[None]
        posevec = fr.dst_posevec[None]
        # network.py:558-571 with the reference's own modules / static method
        refined = ref.pose_decoder(posevec)["Rs"]
        no_root = ref._multiply_corrected_Rs(dst_Rs[:, 1:, ...], refined)
        dst_Rs_ref = torch.cat([dst_Rs[:, 0:1, ...], no_root], dim=1)
        for tag, R in (("plain", dst_Rs), ("refined", dst_Rs_ref)):
            mRs, mTs = ref._get_motion_base(dst_Rs=R, dst_Ts=dst_Ts, cnl_gtfms=gt)          # :588-591
            out[f"motion_scale_Rs_{tag}"], out[f"motion_Ts_{tag}"] = mRs.numpy(), mTs.numpy()
        vol = ref.mweight_vol_decoder(motion_weights_priors=sub.priors[None])[0]              # :592-594
        out["refined_Rs"] = refined.numpy()
        out["vol_shape"] = np.array(vol.shape)
        out["vol_sub"] = vol[:, ::4, ::4, ::4].contiguous().numpy()
        out["vol_channel_sums"] = vol.double().sum(dim=(1, 2, 3)).numpy()
    out["dst_Rs"], out["dst_posevec"] = fr.dst_Rs.numpy(), fr.dst_posevec.numpy()
    out["param_checksum"] = np.array(sum(float(p.detach().double().sum()) for p in pro.parameters()))
    np.savez_compressed(PATH, **out)
    print(PATH, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
