"""GPU-side "beat THAT kernel" table for the hash grid: the reference's own
gridencoder.cu (compiled unmodified for sm_100a by oracle/build_ref.sh -> oracle/_ref/_gridencoder_ref.so) timed next to
occnerf_hashgrid_forward / _backward on the SAME inputs, B = 786 432 and 2^20 .. 2^22, D=4, L=16, C=2, table 2^19.

Inputs: `ray` = encoder inputs of the bench workload's samples (ordered along rays, what the path sees; tiled to size B),
`uniform` = i.i.d. uniform points (what an unordered caller would pass).  Timing: CUDA events over `reps` launches after
3 warm-ups, a 256 MiB L2 flush before every launch, reference timed with its own output layout [L,B,C] and its own
caller-side costs EXCLUDED (the zeros_like of the 59 MiB gradient and the [L,B,C]->[B,L*C] permute copies of grid.py:58,76
are listed separately).  Writes a markdown table to stdout / gpurun_out/hashgrid_vs_ref.md."""
import importlib.util, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from occnerf_b200 import _lib, ops, synthetic as S
from occnerf_b200.network import RenderConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "_gridencoder_ref.so")
spec = importlib.util.spec_from_file_location("_gridencoder_ref", REF_SO)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)
d = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024 // 4, device=d)


def timed(fn, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def ray_inputs():
    """encoder inputs (p, normalised signed distance) of the bench workload's 786 432 samples, in ray order"""
    sub = S.make_subject(seed=0)
    w = S.make_weights(sub.bound, seed=0)
    net = S.network_from_synthetic(sub, w, RenderConfig(perturb=0.0, mlp_engine="tf32"), device=d).train(False)
    fr = S.frame_to(S.make_frame(sub, mode="patch", n_patches=6, patch=32, seed=100), d)
    vol = S.make_motion_weights_vol(sub.priors, seed=0).to(d)
    packed = torch.cat([fr.rays_o, fr.rays_d, fr.near, fr.far], -1).contiguous()
    z, x_skel, mask = ops.warp_forward(packed, None, fr.motion_scale_Rs.contiguous(), fr.motion_Ts.contiguous(), vol, fr.cnl_bbox_min_xyz.contiguous(),
                                       fr.cnl_bbox_scale_xyz.contiguous(), 128)
    xyz = x_skel.reshape(-1, 3).contiguous()
    knn = net._knn(xyz, 128)
    st = net._static()
    enc_in, _ = ops.sample_geometry(xyz, knn, st["point_base"], st["point_norms"], net.bound)
    enc = net.cnl_mlp.module.encoder
    return enc_in, enc


enc_ray, enc = ray_inputs()
emb = (torch.rand(enc.embeddings.shape, device=d) * 2 - 1) * 1e-4
offs = enc.offsets.to(d)
Sv = float(np.log2(enc.per_level_scale))
scales = ops.level_scales(Sv, 16, 16, d)
rows = []
for B in (786432, 1 << 20, 1 << 21, 1 << 22):
    for kind in ("ray", "uniform"):
        if kind == "ray":
            x = enc_ray.repeat((B + enc_ray.shape[0] - 1) // enc_ray.shape[0], 1)[:B].contiguous()
        else:
            x = torch.rand(B, 4, device=d)
        out_ref = torch.empty(16, B, 2, device=d)
        g_lbc = torch.randn(16, B, 2, device=d)
        g_blc = g_lbc.permute(1, 0, 2).reshape(B, 32).contiguous()
        ge_ref, ge = torch.zeros_like(emb), torch.zeros_like(emb)
        out = torch.empty(B, 32, device=d)
        t = {}
        t["ref_fwd"] = timed(lambda: ref.grid_encode_forward(x, emb, offs, out_ref, B, 4, 2, 16, Sv, 16, None, 0, False, 0))
        t["ref_bwd"] = timed(lambda: ref.grid_encode_backward(g_lbc, x, emb, offs, ge_ref, B, 4, 2, 16, Sv, 16, None, None, 0, False, 0))
        t["ref_caller_copies"] = timed(lambda: (out_ref.permute(1, 0, 2).reshape(B, 32).contiguous(), g_blc.view(B, 16, 2).permute(1, 0, 2).contiguous(),
                                                torch.zeros_like(emb)))
        for rl, tag in ((0, "per_sample"), (ops.HASH_BWD_RUN, "runs")):
            if kind == "uniform" and rl:
                continue
            t[f"ours_fwd_{tag}"] = timed(lambda: ops.hashgrid_forward(x, emb, offs, scales, out_ptr=out.data_ptr(), ld=32, run_length=rl))
            t[f"ours_bwd_{tag}"] = timed(lambda: ops.hashgrid_backward(g_blc.data_ptr(), 32, 0, x, offs, scales, ge, 2, run_length=rl))
        # same values (forward bitwise, gradient to summation order)
        ops.hashgrid_forward(x, emb, offs, scales, out_ptr=out.data_ptr(), ld=32, run_length=0)
        ref.grid_encode_forward(x, emb, offs, out_ref, B, 4, 2, 16, Sv, 16, None, 0, False, 0)
        same = bool(torch.equal(out, out_ref.permute(1, 0, 2).reshape(B, 32)))
        rows.append(dict(B=B, inputs=kind, forward_bitwise_equal=same, **{k: round(v, 4) for k, v in t.items()}))
        del x, out_ref, g_lbc, g_blc, out
        torch.cuda.empty_cache()

lines = ["| B | inputs | ref fwd ms | ours fwd ms (per-sample / runs) | ref bwd ms | ours bwd ms (per-sample / runs) | ref caller copies ms | fwd bitwise equal |",
         "|---:|---|---:|---:|---:|---:|---:|---|"]
for r in rows:
    lines.append(f"| {r['B']} | {r['inputs']} | {r['ref_fwd']} | {r['ours_fwd_per_sample']} / {r.get('ours_fwd_runs', '-')} | {r['ref_bwd']} | "
                 f"{r['ours_bwd_per_sample']} / {r.get('ours_bwd_runs', '-')} | {r['ref_caller_copies']} | {r['forward_bitwise_equal']} |")
md = "\n".join(lines)
print(md)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "hashgrid_vs_ref.md"), "w").write(md + "\n\n" + json.dumps(rows) + "\n")
