"""Sharded MSM on two real GPUs (skipped on a single-GPU box): one process per GPU, exchange over NVLink peer-memory
mailboxes and over NCCL, then with prepared points and with fixed-base tables, each against this rank's own single-GPU
MSM; identical bits on both ranks."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)     # plumbing only: handles / ids travel over it
    import dusk_zerocaf_b200 as zc
    from dusk_zerocaf_b200 import synth
    ctx = zc.Context(rank)
    L = ctx._L
    n, c = 50_000, 16
    dev = torch.device("cuda", rank)
    sc = torch.from_numpy(synth.synth_scalar(300, 0, n).view(np.int64)).to(dev)
    P = torch.empty((n, 20), dtype=torch.int64, device=dev)
    ctx.check(L.zc_basepoint_mul_batch_dev(ctx._h, sc.data_ptr(), P.data_ptr(), n))
    S = torch.from_numpy(synth.synth_scalar(301, 0, n).view(np.int64)).to(dev)
    full = torch.zeros(20, dtype=torch.int64, device=dev)
    ctx.check(L.zc_msm_dev(ctx._h, P.data_ptr(), S.data_ptr(), n, c, full.data_ptr()))

    def allgather(b):
        objs = [None] * world
        dist.all_gather_object(objs, b)
        return objs

    def bcast(b):
        obj = [b]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]

    results = []
    ctx.init_nccl(rank, world, bcast)
    gens = None
    for path in ("nccl", "peer", "peer+prepared", "peer+fixed_base"):
        if path == "peer":
            ctx.init_peer_mailboxes(rank, world, allgather)
        elif path == "peer+prepared":
            gens = ctx.msm_generators(P.data_ptr(), n, zc.GEN_PREPARED)
        elif path == "peer+fixed_base":
            gens.close()
            gens = ctx.msm_generators(P.data_ptr(), n, zc.GEN_FIXED_BASE, c, rank, world)
        out = torch.zeros(20, dtype=torch.int64, device=dev)
        for _ in range(3):                                               # repeated calls: sequence numbers / graph replay
            if gens is not None:
                gens.msm_sharded(S.data_ptr(), out.data_ptr(), window_bits=c)
            else:
                ctx.check(L.zc_msm_sharded_dev(ctx._h, P.data_ptr(), S.data_ptr(), n, c, out.data_ptr()))
        eq = torch.zeros(1, dtype=torch.uint8, device=dev)
        ctx.check(L.zc_ristretto_eq_batch_dev(ctx._h, out.data_ptr(), full.data_ptr(), eq.data_ptr(), 1))
        ctx.sync()
        results.append((path, bool(eq.item()), out.cpu().numpy().tobytes()))
    q.put((rank, results))
    gens.close()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_msm_two_gpus_peer_and_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    world = 2
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for r in range(world):
        for path, ok, _ in res[r]:
            assert ok, (r, path)
    for k in range(len(res[0])):
        assert res[0][k][2] == res[1][k][2], res[0][k][0]            # identical bits on all ranks, per path
