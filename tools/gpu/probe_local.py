"""Where does a single host thread block when it enqueues the ranks of a sharded MSM one after the other? (debug aid)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import dusk_zerocaf_b200 as zc
from dusk_zerocaf_b200 import synth
R, K, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
streams = [torch.cuda.Stream() for _ in range(R)]
ctxs = [zc.Context(0, stream=s.cuda_stream) for s in streams]
zc.Context.connect_local(ctxs)
sc = torch.from_numpy(synth.synth_scalar(100, 0, n).view(np.int64)).cuda()
P = torch.empty((n, 20), dtype=torch.int64, device="cuda")
ctxs[0].check(ctxs[0]._L.zc_basepoint_mul_batch_dev(ctxs[0]._h, sc.data_ptr(), P.data_ptr(), n))
ctxs[0].sync()
S = [torch.from_numpy(synth.synth_scalar(101 + j, 0, n).view(np.int64)).cuda() for j in range(2)]
out = torch.zeros((R, K, 20), dtype=torch.int64, device="cuda")
warm = torch.zeros(20, dtype=torch.int64, device="cuda")
for r, cx in enumerate(ctxs):
    cx.check(cx._L.zc_msm_partial_dev(cx._h, P.data_ptr(), S[0].data_ptr(), n, 16, r, R, warm.data_ptr()))
    cx.sync()
t00 = time.perf_counter()
for k in range(K):
    for r, cx in enumerate(ctxs):
        t0 = time.perf_counter()
        cx.check(cx._L.zc_msm_sharded_dev(cx._h, P.data_ptr(), S[k % 2].data_ptr(), n, 16, out[r, k].data_ptr()))
        print(f"call {k} rank {r}: host {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
for r, cx in enumerate(ctxs):
    t0 = time.perf_counter()
    try:
        cx.sync()
        print(f"sync rank {r}: {1e3 * (time.perf_counter() - t0):.2f} ms ok")
    except Exception as e:
        print(f"sync rank {r}: {1e3 * (time.perf_counter() - t0):.2f} ms {e}")
print(f"total {1e3 * (time.perf_counter() - t00):.1f} ms; ranks agree: {bool(torch.equal(out[0], out[R - 1]))}")
