"""CPU: why single entries of d loss / d point_dist (and of the table gradient) get a looser tolerance than everything else.

The hash grid is piecewise multilinear: its derivative w.r.t. the INPUT jumps at cell boundaries.  The 6890 per-vertex inputs
(network.py:263-284 -> occnerf_mlp.py:171-175) are fixed points, and a few of them sit within one ulp of a boundary of some level,
where `floor(x * scale_l + 0.5)` depends on the last bit of `scale_l = exp2f(l * S) * 16 - 1` -- a value the reference computes with
the DEVICE's exp2f (gridencoder.cu:138) and a host restatement computes with libm.  This test moves ONE level's scale by one ulp in
the oracle and shows what that does: outputs and the overall gradient barely move, but individual vertices' gradients flip to the
derivative of the neighbouring cell.  (Found with tools/diag_point_dist.py: vertex 4963 of the golden case sits on a level-14
boundary; left derivative +10.0, right derivative -2.8, and the whole 1.9e-2 per-entry difference round 1 reported came from it.)
With the device's own scale table handed to the oracle the difference disappears: tests/test_render_gpu.py holds that comparison to 2e-3."""
import copy

import numpy as np
import torch

from oracle import hashgrid_c, make_golden, occnerf_oracle as O
from tests.helpers import load_case


def test_one_ulp_of_a_level_scale_flips_single_vertex_gradients():
    sub, w, fr, vol, t_rand, rk, g = load_case("train_dense")
    S = float(np.log2(w.per_level_scale))
    host = hashgrid_c.host_level_scales(S, 16, 16)

    def run(scales):
        s, ww = copy.deepcopy(sub), copy.deepcopy(w)
        for t in [ww.embeddings, s.point_dist]:
            t.requires_grad_(True)
        o = O.render_rays(fr, vol, s, ww, iter_val=rk["iter_val"], training=True, t_rand=t_rand, level_scales=scales)
        make_golden.scalar_loss(o).backward()
        return o, s.point_dist.grad.reshape(-1).double(), ww.embeddings.grad.double()

    o0, pd0, e0 = run(host.clone())
    bumped = host.clone()
    bumped[14] = torch.nextafter(bumped[14], torch.tensor(1e9))
    o1, pd1, e1 = run(bumped)
    out_diff = max(float((o0[k] - o1[k]).abs().max()) for k in ("rgb", "alpha", "depth"))
    per_entry = float((pd1 - pd0).abs().max() / pd0.abs().max())
    fro = float((pd1 - pd0).norm() / pd0.norm())
    n_moved = int(((pd1 - pd0).abs() > 1e-3 * pd0.abs().max()).sum())
    assert out_diff < 1e-5                               # the rendered values do not care
    assert per_entry > 5e-3, per_entry                   # ... single vertices do: this is the sensitivity the 2e-2 tolerance covers
    assert fro < 2e-2 and n_moved < 20, (fro, n_moved)   # ... and they are few
    assert float((e1 - e0).norm() / e0.norm()) < 5e-3
