"""Multi-GPU MSM decomposition on CPU: world_size-2 gloo processes model the bucket-window sharding with the oracle as
the group (each rank sums ITS windows, one all_gather of the 160-byte partial points, fixed-order fold) and must agree
with the naive MSM and with each other bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import SEED
from dusk_zerocaf_b200 import sharding, synth
import msm_plan_model as model


def test_window_plan():
    assert sharding.num_windows(16) == 16
    owned = [sharding.windows_of_rank(16, r, 8) for r in range(8)]
    assert sorted(sum(owned, [])) == list(range(16)) and all(len(o) == 2 for o in owned)
    assert sharding.windows_of_rank(16, 3, 32 if False else 16) == [3]
    L = 2**249 + 14490550575682688738086195780655237219
    rng = np.random.default_rng(1)
    for c in range(8, 17):
        for s in [0, 1, L - 1, 2**249 - 1, 2**(c - 1), 2**c - 1] + [int(rng.integers(0, 2**62)) << 180 for _ in range(20)]:
            s %= L
            d = model.signed_digits(s, c)
            assert sum(x << (c * w) for w, x in enumerate(d)) == s
            assert all(-(1 << (c - 1)) <= x < (1 << (c - 1)) for x in d)


def test_task_plan_covers_every_window_point_once():
    for c, R, n in [(16, 8, 10), (16, 8, 7), (16, 4, 5), (16, 2, 5), (16, 1, 3), (8, 16, 6), (13, 8, 4), (16, 16, 4)]:
        seen = {}
        for r in range(R):
            t = sharding.tasks_of_rank(c, r, R, n)
            assert [w for w, _, _ in t] == sorted(w for w, _, _ in t)
            for w, p0, p1 in t:
                for i in range(p0, p1):
                    assert (w, i) not in seen, (c, R, w, i)
                    seen[(w, i)] = r
        assert len(seen) == sharding.num_windows(c) * n
    assert all(len(sharding.tasks_of_rank(16, r, 8, 100)) == 2 for r in range(8))


def _model_partial(o, P, S, c, rank, world):
    """One rank's partial point: Horner over all windows, adding only the (window, point range) tasks it owns."""
    n = P.shape[0]
    digs = [model.signed_digits(model.limbs_to_int(S[i]), c) for i in range(n)]
    nwin = sharding.num_windows(c)
    mine = {w: (p0, p1) for w, p0, p1 in sharding.tasks_of_rank(c, rank, world, n)}
    acc, started = o.pt_identity(), False
    for w in range(nwin - 1, -1, -1):
        if started:
            for _ in range(c):
                acc = o.pt_double(acc)
        if w in mine:
            ws = o.pt_identity()
            for i in range(*mine[w]):
                d = digs[i][w]
                if d:
                    t = o.pt_double_and_add(P[i], o.int_to_limbs(abs(d)))
                    ws = o.pt_add(ws, o.pt_neg(t) if d < 0 else t)
            acc = o.pt_add(acc, ws)
            started = True
    return acc


def test_plan_model_8_ranks(oracle):
    """The 8-rank plan folds to the naive MSM -- the oracle as the group."""
    o = oracle
    n, c, world = 5, 16, 8
    P = o.pt_scalar_mul_batch(np.tile(synth.BASEPOINT, (n, 1)), synth.synth_scalar(210, 0, n))
    S = synth.synth_scalar(211, 0, n)
    total = _model_partial(o, P, S, c, 0, world)
    for r in range(1, world):
        total = o.pt_add(total, _model_partial(o, P, S, c, r, world))
    assert o.pt_eq(total, o.msm_naive(P, S))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, c, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as o
    base = np.tile(synth.BASEPOINT, (n, 1))
    P = o.pt_scalar_mul_batch(base, synth.synth_scalar(200, 0, n))
    S = synth.synth_scalar(201, 0, n)
    acc = _model_partial(o, P, S, c, rank, world)
    part = torch.from_numpy(acc.view(np.int64).copy())
    gathered = [torch.zeros(20, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, part)                          # the ONE exchange step
    total = gathered[0].numpy().view(np.uint64).copy()
    for r in range(1, world):                                # fixed-order fold, edwards.rs:465-489
        total = o.pt_add(total, gathered[r].numpy().view(np.uint64))
    want = o.msm_naive(P, S)
    q.put((rank, bool(o.pt_eq(total, want)), total.tobytes()))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_gloo_world2_sharded_msm_model():
    world, n, c = 2, 6, 16
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, c, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == res[1][2]          # identical bits on all ranks
