"""Checks and times occnerf_allreduce_sum_f32 (SwitchReducer) against NCCL under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/allreduce_check.py
Sizes are those of the training step: the 59.2 MiB hash-table gradient + the small-gradient bucket."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from occnerf_b200.distributed import SwitchReducer

TABLE, BUCKET = 7755336 * 2, 1 << 20
res = {"world": world}
red = SwitchReducer(TABLE, BUCKET, dev)
res["kind"] = red.kind
res["multicast_ptr_nonzero"] = bool(red.multicast)


def fill(seed):
    g = torch.Generator(device=dev).manual_seed(seed * 100 + rank)
    red.flat.copy_(torch.randn(red.flat.numel(), device=dev, generator=g))


def check(tag, force_p2p=False):
    fill(1)
    small = [torch.randn(1000, 37, device=dev) + rank, torch.ones(6890, device=dev) * (rank % 2)]
    ref_flat = red.flat[:TABLE].clone()
    ref_small = [t.clone() for t in small]
    dist.all_reduce(ref_flat)
    for t in ref_small:
        dist.all_reduce(t)
    keep = red.multicast
    if force_p2p:
        red.multicast = 0
    red([red.table_view] + [small[0]], hits=small[1])
    red.multicast = keep
    torch.cuda.synchronize()
    e_tab = float((red.flat[:TABLE] - ref_flat).abs().max() / ref_flat.abs().max())
    e_small = float((small[0] - ref_small[0]).abs().max() / ref_small[0].abs().max())
    hits_ok = bool(torch.equal(small[1], ref_small[1].clamp(max=1.0)))
    res[tag] = {"table_rel_err": e_tab, "small_rel_err": e_small, "hits_ok": hits_ok}


check("multimem" if red.multicast else "p2p")
if red.multicast:
    check("p2p_forced", force_p2p=True)

# timing: the kernel alone over the table gradient vs one NCCL all-reduce of the same bytes


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


res["switch_ms"] = timed(lambda: red([red.table_view]))
res["bytes"] = red.used * 4
x = torch.randn(red.used, device=dev)
if red.multicast:
    keep, red.multicast = red.multicast, 0
    res["p2p_ms"] = timed(lambda: red([red.table_view]))
    red.multicast = keep
res["nccl_ms"] = timed(lambda: dist.all_reduce(x))
# library multimem kernels of torch's symmetric memory on the same buffer, as a yardstick of what the fabric gives
try:
    gname = dist.group.WORLD.group_name
    res["torch_multimem_all_reduce_ms"] = timed(lambda: torch.ops.symm_mem.multimem_all_reduce_(red.flat[:red.used], "sum", gname))
    res["torch_two_shot_all_reduce_ms"] = timed(lambda: torch.ops.symm_mem.two_shot_all_reduce_(red.flat[:red.used], "sum", gname))
except Exception as exc:
    res["torch_symm_mem_ops"] = f"{type(exc).__name__}: {exc}"[:200]
for b in (16, 32, 128):
    r2 = SwitchReducer(TABLE, BUCKET, dev, blocks=b)
    r2.used = red.used
    r2.key = ()
    res[f"switch_ms_blocks{b}"] = timed(lambda: r2([r2.table_view]))

# the same launch right after 4 GB of unrelated traffic (what the training step does before it: cold L2, cold TLBs)
big = torch.empty(1 << 30, device=dev)


def timed_cold(fn, n=10):
    tot = 0.0
    for i in range(n + 2):
        big.add_(1.0)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= 2:
            tot += e0.elapsed_time(e1)
    t = torch.tensor([tot / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


res["switch_ms_after_4GB_traffic"] = timed_cold(lambda: red([red.table_view]))
res["nccl_ms_after_4GB_traffic"] = timed_cold(lambda: dist.all_reduce(x))
try:
    res["torch_multimem_ms_after_4GB_traffic"] = timed_cold(lambda: torch.ops.symm_mem.multimem_all_reduce_(red.flat[:red.used], "sum", gname))
except Exception as exc:
    pass
import ctypes
buf = (ctypes.c_ulonglong * 4)()
from occnerf_b200 import _lib as L
L.load().occnerf_allreduce_debug(ctypes.cast(buf, ctypes.c_void_p), 1)
timed_cold(lambda: red([red.table_view]))
L.load().occnerf_allreduce_debug(ctypes.cast(buf, ctypes.c_void_p), 1)
res["cold_phases_us"] = {"wait_arrive": buf[0] / max(buf[3], 1) / 1e3, "data": buf[1] / max(buf[3], 1) / 1e3, "wait_finish": buf[2] / max(buf[3], 1) / 1e3}
del big
# NVLink idle between launches (the training step leaves the links idle for ~8 ms): an 8 ms matmul burst in the stream before every
# all-reduce, no host sync and no other collective in between
A = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)


def timed_after_compute(fn, n=10, burst=8):
    evs = []
    torch.cuda.synchronize(); dist.barrier()
    for i in range(n + 2):
        for _ in range(burst):
            A @ A
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs[2:]) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


L.load().occnerf_allreduce_debug(ctypes.cast(buf, ctypes.c_void_p), 1)
res["switch_ms_after_8ms_compute"] = timed_after_compute(lambda: red([red.table_view]))
L.load().occnerf_allreduce_debug(ctypes.cast(buf, ctypes.c_void_p), 1)
res["after_compute_phases_us"] = {"wait_arrive": buf[0] / max(buf[3], 1) / 1e3, "data": buf[1] / max(buf[3], 1) / 1e3, "wait_finish": buf[2] / max(buf[3], 1) / 1e3}
res["nccl_ms_after_8ms_compute"] = timed_after_compute(lambda: dist.all_reduce(x))
res["switch_ms_after_1ms_compute"] = timed_after_compute(lambda: red([red.table_view]), burst=1)
del A
# does it matter HOW the buffer was written?  (in the training step: a memset and red.global atomics)
idx = torch.randint(0, TABLE, (8 << 20,), device=dev)
val = torch.randn(8 << 20, device=dev)


def timed_after(prep, fn, n=10):
    evs = []
    torch.cuda.synchronize(); dist.barrier()
    for i in range(n + 2):
        prep()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs[2:]) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def phases():
    L.load().occnerf_allreduce_debug(ctypes.cast(buf, ctypes.c_void_p), 1)
    return {"wait_arrive": buf[0] / max(buf[3], 1) / 1e3, "data": buf[1] / max(buf[3], 1) / 1e3, "wait_finish": buf[2] / max(buf[3], 1) / 1e3}


phases()
res["switch_ms_after_memset"] = timed_after(lambda: red.flat.zero_(), lambda: red([red.table_view]))
res["after_memset_phases_us"] = phases()
res["switch_ms_after_atomics"] = timed_after(lambda: red.flat.index_add_(0, idx, val), lambda: red([red.table_view]))
res["after_atomics_phases_us"] = phases()
res["switch_ms_after_copy"] = timed_after(lambda: red.flat[:TABLE].copy_(x[:TABLE]), lambda: red([red.table_view]))
res["after_copy_phases_us"] = phases()

# CUDA-graph replay: the kernel's epochs live in device memory, so a captured launch keeps working
fill(2)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    red([red.table_view])
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    red([red.table_view])
fill(3)
ref = red.flat[:TABLE].clone()
dist.all_reduce(ref)
g.replay()
torch.cuda.synchronize()
res["graph_replay_rel_err"] = float((red.flat[:TABLE] - ref).abs().max() / ref.abs().max())
fill(4)
ref = red.flat[:TABLE].clone()
dist.all_reduce(ref)
g.replay()
torch.cuda.synchronize()
res["graph_replay2_rel_err"] = float((red.flat[:TABLE] - ref).abs().max() / ref.abs().max())
if rank == 0:
    print(json.dumps(res, indent=1))
dist.barrier()
dist.destroy_process_group()
