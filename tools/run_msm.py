"""Run the 2^20-point MSM (BASELINE config 5) a few times on one GPU -- the target of ncu captures and quick timings."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dusk_zerocaf_b200 as zc
from dusk_zerocaf_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--c", type=int, default=16)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--rank", type=int, default=0)
ap.add_argument("--nranks", type=int, default=1)
ap.add_argument("--check", action="store_true")
ap.add_argument("--prepared", action="store_true")
ap.add_argument("--fixed-base", action="store_true", help="pre-scaled per-window tables (zc_msm_prepare_fixed_base_dev)")
a = ap.parse_args()

dev = torch.device("cuda", 0)
st = torch.cuda.Stream()
ctx = zc.Context(0, stream=st.cuda_stream)
L = ctx._L
n = a.n
sc = torch.from_numpy(synth.synth_scalar(100, 0, n).view(np.int64)).to(dev)
base = torch.from_numpy(np.tile(synth.BASEPOINT, (n, 1)).view(np.int64)).to(dev)
P = torch.empty((n, 20), dtype=torch.int64, device=dev)
ctx.check(L.zc_point_scalar_mul_batch_dev(ctx._h, base.data_ptr(), sc.data_ptr(), P.data_ptr(), n, 1))
S = torch.from_numpy(synth.synth_scalar(102, 0, n).view(np.int64)).to(dev)
out = torch.zeros(20, dtype=torch.int64, device=dev)
ctx.sync()
gens = None
if a.prepared:
    gens = ctx.msm_generators(P.data_ptr(), n, zc.GEN_PREPARED)
if a.fixed_base:
    t0 = time.perf_counter()
    gens = ctx.msm_generators(P.data_ptr(), n, zc.GEN_FIXED_BASE, a.c, a.rank, a.nranks)
    print(f"fixed-base tables rank {a.rank}/{a.nranks}: {(time.perf_counter() - t0) * 1e3:.1f} ms, {gens.device_bytes >> 20} MiB", flush=True)
for it in range(a.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    if gens is not None:
        gens.msm_partial(S.data_ptr(), out.data_ptr(), a.rank, a.nranks, window_bits=a.c)
    else:
        ctx.check(L.zc_msm_partial_dev(ctx._h, P.data_ptr(), S.data_ptr(), n, a.c, a.rank, a.nranks, out.data_ptr()))
    e1.record(st)
    st.synchronize()
    print(f"msm n={n} c={a.c} rank {a.rank}/{a.nranks}: {e0.elapsed_time(e1):.3f} ms", flush=True)
if a.check and a.fixed_base:
    # the same MSM through the plain path (no tables) must be the same group element
    ctx.sync()
    ref = torch.zeros(20, dtype=torch.int64, device=dev)
    ctx.check(L.zc_msm_partial_dev(ctx._h, P.data_ptr(), S.data_ptr(), n, a.c, a.rank, a.nranks, ref.data_ptr()))
    ctx.sync()
    eq = zc.batch.ristretto_eq(out.cpu().numpy().view(np.uint64)[None], ref.cpu().numpy().view(np.uint64)[None])
    print("fixed-base partial == plain partial:", bool(eq[0]))
if a.check:
    # first m points: the MSM against the sum of the strict scalar multiplications, folded on the device (no CPU code here;
    # the oracle comparisons live in tests/)
    m = min(n, 2048)
    got = torch.zeros(20, dtype=torch.int64, device=dev)
    ctx.check(L.zc_msm_dev(ctx._h, P.data_ptr(), S.data_ptr(), m, a.c, got.data_ptr()))
    prods = torch.empty((m, 20), dtype=torch.int64, device=dev)
    ctx.check(L.zc_point_scalar_mul_batch_dev(ctx._h, P.data_ptr(), S.data_ptr(), prods.data_ptr(), m, 0))
    want = torch.zeros(20, dtype=torch.int64, device=dev)
    ctx.check(L.zc_point_fold_dev(ctx._h, prods.data_ptr(), m, want.data_ptr()))
    eq = torch.zeros(1, dtype=torch.uint8, device=dev)
    ctx.check(L.zc_ristretto_eq_batch_dev(ctx._h, got.data_ptr(), want.data_ptr(), eq.data_ptr(), 1))
    ctx.sync()
    print("check first", m, "points:", bool(eq.item()))
