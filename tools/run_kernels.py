"""Launch each hot kernel a few times on one GPU -- the target of the ncu captures under profiles/ (round 2)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dusk_zerocaf_b200 as zc
from dusk_zerocaf_b200 import synth

dev = torch.device("cuda", 0)
st = torch.cuda.Stream()
ctx = zc.Context(0, stream=st.cuda_stream)
L = ctx._L


def dev_u64(a):
    return torch.from_numpy(a.view(np.int64)).to(dev)


n2 = 1 << 24
a, b = dev_u64(synth.synth_fe(1, 0, n2)), dev_u64(synth.synth_fe(2, 0, n2))
p, q = torch.empty_like(a), torch.empty_like(a)
for _ in range(2):
    ctx.check(L.zc_fe_mul_square_batch_dev(ctx._h, a.data_ptr(), b.data_ptr(), p.data_ptr(), q.data_ptr(), n2))
# the same on the 32-byte wire format
ab = torch.empty((n2, 32), dtype=torch.uint8, device=dev)
bb, pb, qb = torch.empty_like(ab), torch.empty_like(ab), torch.empty_like(ab)
ctx.check(L.zc_fe_to_bytes_batch_dev(ctx._h, a.data_ptr(), ab.data_ptr(), n2))
ctx.check(L.zc_fe_to_bytes_batch_dev(ctx._h, b.data_ptr(), bb.data_ptr(), n2))
for _ in range(2):
    ctx.check(L.zc_fe_mul_square_batch_packed_dev(ctx._h, ab.data_ptr(), bb.data_ptr(), pb.data_ptr(), qb.data_ptr(), n2))
ctx.sync()
del a, b, p, q, ab, bb, pb, qb
n3 = 1 << 22
s1, s2 = dev_u64(synth.synth_scalar(100, 0, n3)), dev_u64(synth.synth_scalar(101, 0, n3))
P, Q, O = (torch.empty((n3, 20), dtype=torch.int64, device=dev) for _ in range(3))
ctx.check(L.zc_basepoint_mul_batch_dev(ctx._h, s1.data_ptr(), P.data_ptr(), n3))
ctx.check(L.zc_basepoint_mul_batch_dev(ctx._h, s2.data_ptr(), Q.data_ptr(), n3))
for _ in range(2):
    ctx.check(L.zc_point_add_batch_dev(ctx._h, P.data_ptr(), Q.data_ptr(), O.data_ptr(), n3))
n4 = 1 << 18
for mode in (0, 1, 0, 1):
    ctx.check(L.zc_point_scalar_mul_batch_dev(ctx._h, P.data_ptr(), s2.data_ptr(), O.data_ptr(), n4, mode))
n5 = 1 << 20
out = torch.zeros(20, dtype=torch.int64, device=dev)
g = ctx.msm_generators(P.data_ptr(), n5, zc.GEN_PREPARED)
for _ in range(2):
    g.msm(s2.data_ptr(), out.data_ptr(), window_bits=16)
ctx.sync()
g.close()
print("done", ctx.launches, "launches")
