"""Time single kernels through the C ABI on one GPU (CUDA events on the context's stream): A/B runs of kernel variants
selected by environment variables (ZC_PT_VARIANT, ZC_FIXED_LDG, ...).  usage: python tools/time_ops.py [pt] [fixed] [smul]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dusk_zerocaf_b200 as zc
from dusk_zerocaf_b200 import synth

what = set(sys.argv[1:]) or {"pt", "fixed"}
dev = torch.device("cuda", 0)
st = torch.cuda.Stream()
ctx = zc.Context(0, stream=st.cuda_stream)
L = ctx._L


def dev_u64(a):
    return torch.from_numpy(a.view(np.int64)).to(dev)


def timed(fn, k=10, w=3):
    for _ in range(w):
        fn()
    st.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(k):
        fn()
    e1.record(st)
    st.synchronize()
    return e0.elapsed_time(e1) / k


tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("ZC_"))
n3 = 1 << 22
s1, s2 = dev_u64(synth.synth_scalar(100, 0, n3)), dev_u64(synth.synth_scalar(101, 0, n3))
P, Q, O = (torch.empty((n3, 20), dtype=torch.int64, device=dev) for _ in range(3))
ctx.check(L.zc_basepoint_mul_batch_dev(ctx._h, s1.data_ptr(), P.data_ptr(), n3))
ctx.check(L.zc_basepoint_mul_batch_dev(ctx._h, s2.data_ptr(), Q.data_ptr(), n3))
ctx.sync()
if "pt" in what:
    ms = timed(lambda: ctx.check(L.zc_point_add_batch_dev(ctx._h, P.data_ptr(), Q.data_ptr(), O.data_ptr(), n3)))
    msd = timed(lambda: ctx.check(L.zc_point_double_batch_dev(ctx._h, P.data_ptr(), O.data_ptr(), n3)))
    print(f"[{tag}] point add 2^22: {ms:.4f} ms ({n3 / ms / 1e6:.3f} G/s)   double: {msd:.4f} ms")
if "fixed" in what:
    n4 = 1 << 20
    ms = timed(lambda: ctx.check(L.zc_basepoint_mul_batch_dev(ctx._h, s1.data_ptr(), O.data_ptr(), n4)), k=5)
    print(f"[{tag}] basepoint mul 2^20: {ms:.4f} ms ({n4 / ms / 1e3:.1f} M/s)")
if "smul" in what:
    n4 = 1 << 20
    for mode in (0, 1):
        ms = timed(lambda: ctx.check(L.zc_point_scalar_mul_batch_dev(ctx._h, P.data_ptr(), s2.data_ptr(), O.data_ptr(), n4, mode)), k=2, w=1)
        print(f"[{tag}] scalar mul mode {mode} 2^20: {ms:.3f} ms")
ctx.close()
