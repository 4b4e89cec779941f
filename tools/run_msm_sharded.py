"""The 2^20-point MSM (BASELINE config 5) sharded by bucket-window over the ranks of one node -- only that, so a multi-GPU
box is held for seconds: `python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/run_msm_sharded.py`
(or plain `python tools/run_msm_sharded.py` for one GPU).  Prints one JSON line on rank 0: ms per MSM (max over ranks,
CUDA events on the launching stream) for the plain path, prepared points and fixed-base tables, and the result checks."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dusk_zerocaf_b200 as zc
from dusk_zerocaf_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--c", type=int, default=16)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
st = torch.cuda.Stream()
ctx = zc.Context(local, stream=st.cuda_stream)
L, n, c = ctx._L, a.n, a.c


def barrier():
    if world > 1:
        dist.barrier()


def max_over_ranks(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    st.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        fn()
    e1.record(st)
    st.synchronize()
    barrier()
    return max_over_ranks(e0.elapsed_time(e1)) / steps


sc = torch.from_numpy(synth.synth_scalar(100, 0, n).view(np.int64)).to(dev)
base = torch.from_numpy(np.tile(synth.BASEPOINT, (n, 1)).view(np.int64)).to(dev)
P = torch.empty((n, 20), dtype=torch.int64, device=dev)
ctx.check(L.zc_point_scalar_mul_batch_dev(ctx._h, base.data_ptr(), sc.data_ptr(), P.data_ptr(), n, 1))
S = torch.from_numpy(synth.synth_scalar(102, 0, n).view(np.int64)).to(dev)
out = torch.zeros(20, dtype=torch.int64, device=dev)
ctx.sync()

res = {"n_points": n, "window_bits": c, "n_gpus": world}
if world > 1:
    def bcast(b):
        obj = [b]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]

    def allgather(b):
        objs = [None] * world
        dist.all_gather_object(objs, b)
        return objs
    ctx.init_nccl(rank, world, bcast)
    fn = lambda: ctx.check(L.zc_msm_sharded_dev(ctx._h, P.data_ptr(), S.data_ptr(), n, c, out.data_ptr()))
    res["ms_plain_nccl_exchange"] = timed(fn, 5)
    ctx.init_peer_mailboxes(rank, world, allgather)
else:
    fn = lambda: ctx.check(L.zc_msm_dev(ctx._h, P.data_ptr(), S.data_ptr(), n, c, out.data_ptr()))
res["ms_plain"] = timed(fn, a.iters)
gens = ctx.msm_generators(P.data_ptr(), n, zc.GEN_PREPARED)
if world > 1:
    fn = lambda: gens.msm_sharded(S.data_ptr(), out.data_ptr(), window_bits=c)
else:
    fn = lambda: gens.msm(S.data_ptr(), out.data_ptr(), window_bits=c)
res["ms_prepared_points"] = timed(fn, a.iters)
prepared_pt = out.clone()
if world > 1:
    # new scalars for every MSM, resident on rank 0: one NCCL broadcast of the limb array in front of every call
    def fn_bcast():
        with torch.cuda.stream(st):
            dist.broadcast(S, src=0)
        fn()
    res["ms_prepared_points_incl_scalar_broadcast"] = timed(fn_bcast, a.iters, warmup=2)
gens.close()
gens = ctx.msm_generators(P.data_ptr(), n, zc.GEN_FIXED_BASE, c, rank, world)
res["ms_fixed_base_tables"] = timed(fn, a.iters)
fb_pt = out.clone()

# checks: all ranks hold identical bits; both modes equal this rank's own single-GPU plain MSM as group elements
gens.close()
full = torch.zeros(20, dtype=torch.int64, device=dev)
ctx.check(L.zc_msm_dev(ctx._h, P.data_ptr(), S.data_ptr(), n, c, full.data_ptr()))
eq = torch.zeros(2, dtype=torch.uint8, device=dev)
got, ref = torch.stack([prepared_pt, fb_pt]), torch.stack([full, full])
ctx.check(L.zc_ristretto_eq_batch_dev(ctx._h, got.data_ptr(), ref.data_ptr(), eq.data_ptr(), 2))
ctx.sync()
ok = bool(eq.all().item())
identical = True
if world > 1:
    for pt in (prepared_pt, fb_pt):
        g = [torch.zeros_like(pt) for _ in range(world)]
        dist.all_gather(g, pt)
        identical = identical and all(bool(torch.equal(g[0], x)) for x in g)
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ok = bool(t.item())
res["matches_single_gpu_msm_on_every_rank"] = ok
res["all_ranks_identical_bits"] = identical
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
