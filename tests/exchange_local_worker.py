"""Worker of tests/test_gpu_exchange_local.py: R contexts of ONE process on ONE GPU act as the ranks of the bucket-window-
sharded MSM (zc_peer_mailbox_connect_local), so the fused peer-store / flag / fold exchange kernel -- the code that runs
between GPUs over NVLink -- is exercised and compared with the oracle on a single-GPU box.

Run in its own process with CUDA_DEVICE_MAX_CONNECTIONS=32: every context owns ~10 streams and a spinning exchange kernel
must not share a hardware queue with the work of a rank it waits for.

    python tests/exchange_local_worker.py stress R K n      K back-to-back sharded MSMs of n points on R ranks
    python tests/exchange_local_worker.py timeout           a rank that never arrives becomes ZC_ERR_STATE, not a hang
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SEED = 0x5A45524F43414621


def make_ranks(R):
    import torch
    import dusk_zerocaf_b200 as zc
    streams = [torch.cuda.Stream() for _ in range(R)]
    ctxs = [zc.Context(0, stream=s.cuda_stream) for s in streams]
    zc.Context.connect_local(ctxs)
    return streams, ctxs


def stress(R, K, n):
    import torch
    import dusk_zerocaf_b200 as zc
    from oracle import oracle as o
    o.build()
    from dusk_zerocaf_b200 import synth
    base = np.tile(synth.BASEPOINT, (n, 1))
    P = o.pt_scalar_mul_batch(base, o.synth_scalar(SEED, 400, 0, n), threads=8)
    svec = [o.synth_scalar(SEED, 401 + j, 0, n) for j in range(3)]
    want = [o.msm_naive(P, s, threads=8) for s in svec]
    dP = torch.from_numpy(P.view(np.int64)).cuda()
    dS = [torch.from_numpy(s.view(np.int64)).cuda() for s in svec]
    streams, ctxs = make_ranks(R)
    out = torch.zeros((R, K, 20), dtype=torch.int64, device="cuda")
    warm = torch.zeros(20, dtype=torch.int64, device="cuda")
    # Workspace allocation and the first graph instantiation synchronise the device: do both (each rank's own share of the
    # same shape) before any rank spins in an exchange.  Later calls with other scalars only update the graph in place.
    for r, cx in enumerate(ctxs):
        cx.check(cx._L.zc_msm_partial_dev(cx._h, dP.data_ptr(), dS[0].data_ptr(), n, 16, r, R, warm.data_ptr()))
        cx.sync()
    # scalar set per call: runs of equal sets (graph replay) and alternation (re-capture)
    pick = [(k // 3) % 3 if k % 7 else (k % 3) for k in range(K)]
    for k in range(K):
        for r, cx in enumerate(ctxs):
            cx.check(cx._L.zc_msm_sharded_dev(cx._h, dP.data_ptr(), dS[pick[k]].data_ptr(), n, 16, out[r, k].data_ptr()))
    for cx in ctxs:
        cx.sync()
    res = out.cpu().numpy().view(np.uint64)
    bad_bits = sum(1 for k in range(K) for r in range(1, R) if not np.array_equal(res[0, k], res[r, k]))
    bad_val = sum(1 for k in range(K) if not o.pt_eq(res[0, k], want[pick[k]]))
    # the same ranks through generator handles (prepared points): one more exchange per rank
    gens = [cx.msm_generators(dP.data_ptr(), n, zc.GEN_PREPARED) for cx in ctxs]
    for r, (cx, g) in enumerate(zip(ctxs, gens)):
        g.msm_partial(dS[0].data_ptr(), warm.data_ptr(), r, R, window_bits=16)
        cx.sync()
    out2 = torch.zeros((R, 20), dtype=torch.int64, device="cuda")
    for r, g in enumerate(gens):
        g.msm_sharded(dS[1].data_ptr(), out2[r].data_ptr(), window_bits=16)
    for cx in ctxs:
        cx.sync()
    r2 = out2.cpu().numpy().view(np.uint64)
    gens_ok = all(np.array_equal(r2[0], r2[r]) for r in range(R)) and bool(o.pt_eq(r2[0], want[1]))
    for g in gens:
        g.close()
    launches = sum(cx.launches for cx in ctxs)
    for cx in ctxs:
        cx.close()
    print(json.dumps({"ranks": R, "calls": K, "n": n, "rank_bit_mismatches": bad_bits, "oracle_mismatches": bad_val,
                      "generators_path_ok": gens_ok, "launches": launches}))


def timeout():
    import torch
    import dusk_zerocaf_b200 as zc
    from dusk_zerocaf_b200 import synth
    from oracle import oracle as o
    o.build()
    n = 64
    P = o.pt_scalar_mul_batch(np.tile(synth.BASEPOINT, (n, 1)), o.synth_scalar(SEED, 410, 0, n), threads=4)
    s = o.synth_scalar(SEED, 411, 0, n)
    dP, dS = torch.from_numpy(P.view(np.int64)).cuda(), torch.from_numpy(s.view(np.int64)).cuda()
    streams, ctxs = make_ranks(2)
    out = torch.zeros((2, 20), dtype=torch.int64, device="cuda")
    for r, cx in enumerate(ctxs):
        cx.check(cx._L.zc_msm_partial_dev(cx._h, dP.data_ptr(), dS.data_ptr(), n, 16, r, 2, out[r].data_ptr()))
        cx.sync()
    # only rank 0 calls the collective
    ctxs[0].check(ctxs[0]._L.zc_msm_sharded_dev(ctxs[0]._h, dP.data_ptr(), dS.data_ptr(), n, 16, out[0].data_ptr()))
    status, msg = 0, ""
    try:
        ctxs[0].sync()
    except zc.ZerocafError as e:
        status, msg = e.status, str(e)
    ident = out[0].cpu().numpy().view(np.uint64)
    # reconnect (collective) and run a normal exchange: the ranks are in step again
    zc.Context.connect_local(ctxs)
    for r, cx in enumerate(ctxs):
        cx.check(cx._L.zc_msm_sharded_dev(cx._h, dP.data_ptr(), dS.data_ptr(), n, 16, out[r].data_ptr()))
    for cx in ctxs:
        cx.sync()
    res = out.cpu().numpy().view(np.uint64)
    ok_after = bool(np.array_equal(res[0], res[1]) and o.pt_eq(res[0], o.msm_naive(P, s, threads=4)))
    for cx in ctxs:
        cx.close()
    print(json.dumps({"status": status, "message": msg, "identity_returned": bool(ident[5] == 1 and ident[10] == 1 and ident[0] == 0 and ident[15] == 0),
                      "ok_after_reconnect": ok_after}))


if __name__ == "__main__":
    if sys.argv[1] == "stress":
        stress(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
    else:
        timeout()
