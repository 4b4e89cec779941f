"""Parity at BASELINE.json's full sizes (configs 2-5), through the C ABI.

Where the CPU oracle finishes in seconds (2^24 field products, 2^22 point additions) the comparison is exhaustive and
bit-exact; where it cannot (2^20 scalar multiplications, the 2^20-point MSM) the whole batch is checked through
size-independent properties -- strict vs fast mode agree as group elements, [s]P + [t]P = [s+t]P, MSM = the sum of the
scalar multiplications, MSM is linear in the scalars, the bucket-window shards fold to the full result -- plus a
bit-exact oracle comparison on sampled indices."""
import os

import numpy as np
import pytest

from conftest import SEED

pytestmark = pytest.mark.gpu

THREADS = os.cpu_count() or 8


@pytest.fixture(scope="module")
def zc():
    import dusk_zerocaf_b200 as z
    z.default_context()
    return z


@pytest.fixture(scope="module")
def big_points(zc, oracle):
    """2^22 points [r_i]B with Z != 1 (fast scalar-mul on the device), spot-checked against the oracle."""
    from dusk_zerocaf_b200 import synth
    n = 1 << 22
    r = synth.synth_scalar(100, 0, n)
    P = zc.batch.point_scalar_mul(np.tile(synth.BASEPOINT, (n, 1)), r, mode=1)
    idx = np.concatenate([[0, 1, n - 1], np.random.default_rng(1).integers(0, n, 29)])
    want = oracle.pt_scalar_mul_batch(np.tile(synth.BASEPOINT, (len(idx), 1)), r[idx], threads=THREADS)
    for j, i in enumerate(idx):
        assert oracle.pt_eq(P[i], want[j]), i
    return P


def test_cfg2_field_mul_square_2p24_exhaustive(zc, oracle):
    """config 2: 2^24 pairs, prod = a*b and sq = a^2, every limb against the oracle (field.rs:250-262, 302-315)."""
    from dusk_zerocaf_b200 import synth
    n = 1 << 24
    a, b = synth.synth_fe(1, 0, n), synth.synth_fe(2, 0, n)
    prod, sq = zc.batch.fe_mul_square(a, b)
    want_prod, want_sq = oracle.fe_mul_square_batch(a, b, threads=THREADS)
    assert np.array_equal(prod, want_prod)
    assert np.array_equal(sq, want_sq)
    # the stand-alone kernels agree with the fused one on the whole array
    assert np.array_equal(zc.batch.fe_mul(a, b), prod)
    assert np.array_equal(zc.batch.fe_square(a), sq)


def test_cfg3_point_add_double_2p22_exhaustive(zc, oracle, big_points):
    """config 3: 2^22 additions and doublings, all 20 output limbs against the oracle (edwards.rs:465-489, 579-592)."""
    P = big_points
    Q = np.roll(P, 12345, axis=0)
    assert np.array_equal(zc.batch.point_add(P, Q), oracle.pt_add_batch(P, Q, threads=THREADS))
    assert np.array_equal(zc.batch.point_double(P), oracle.pt_double_batch(P, threads=THREADS))


def test_cfg4_cfg5_scalar_mul_and_msm_2p20(zc, oracle, big_points):
    """configs 4 and 5 at 2^20: properties over the whole batch + sampled bit-exact oracle comparison."""
    from dusk_zerocaf_b200 import synth
    b = zc.batch
    n = 1 << 20
    P = np.ascontiguousarray(big_points[:n])
    s, t = synth.synth_scalar(102, 0, n), synth.synth_scalar(103, 0, n)
    strict = b.point_scalar_mul(P, s, mode=0)
    fast = b.point_scalar_mul(P, s, mode=1)
    # strict mode: limb-exact double_and_add on sampled indices (edwards.rs:102-120)
    idx = np.concatenate([[0, n - 1], np.random.default_rng(2).integers(0, n, 126)])
    assert np.array_equal(strict[idx], oracle.pt_scalar_mul_batch(P[idx], s[idx], threads=THREADS))
    # fast mode: the same group element everywhere (Ristretto equality and encoding)
    assert b.ristretto_eq(strict, fast).all()
    assert np.array_equal(b.ristretto_compress(strict), b.ristretto_compress(fast))
    # [s]P + [t]P = [s + t]P on the whole batch
    st = b.scalar_add(s, t)
    lhs = b.point_add(strict, b.point_scalar_mul(P, t, mode=1))
    assert b.ristretto_eq(lhs, b.point_scalar_mul(P, st, mode=1)).all()

    # config 5: MSM = sum_i [s_i]P_i, the sum taken as a pairwise tree of reference additions over the strict outputs
    acc = strict
    while acc.shape[0] > 1:
        h = acc.shape[0] // 2
        acc = b.point_add(np.ascontiguousarray(acc[:h]), np.ascontiguousarray(acc[h:]))
    msm_s = b.msm(P, s, window_bits=16)
    assert oracle.pt_is_valid(msm_s)
    assert oracle.pt_eq(msm_s, acc[0])
    assert oracle.ris_compress(msm_s) == oracle.ris_compress(acc[0])
    # linear in the scalars: MSM(P, s) + MSM(P, t) = MSM(P, s + t)
    msm_t, msm_st = b.msm(P, t, window_bits=16), b.msm(P, st, window_bits=16)
    assert oracle.pt_eq(oracle.pt_add(msm_s, msm_t), msm_st)
    # other window sizes give the same element
    assert oracle.pt_eq(b.msm(P, s, window_bits=13), msm_s)
    # the bucket-window shards of 8 ranks fold to the full result (the multi-GPU decomposition without NCCL)
    import torch
    ctx = zc.default_context()
    dP = torch.from_numpy(P.view(np.int64)).cuda()
    dS = torch.from_numpy(s.view(np.int64)).cuda()
    parts = torch.zeros((8, 20), dtype=torch.int64, device="cuda")
    for r in range(8):
        ctx.check(ctx._L.zc_msm_partial_dev(ctx._h, dP.data_ptr(), dS.data_ptr(), n, 16, r, 8, parts[r].data_ptr()))
    out = torch.zeros(20, dtype=torch.int64, device="cuda")
    ctx.check(ctx._L.zc_point_fold_dev(ctx._h, parts.data_ptr(), 8, out.data_ptr()))
    ctx.sync()
    assert oracle.pt_eq(out.cpu().numpy().view(np.uint64), msm_s)


@pytest.mark.timeout(600)
def test_cfg5_msm_2p20_against_the_oracle_directly(zc, oracle, big_points):
    """the 2^20-point MSM compared with the ORACLE's naive MSM itself (fold of double_and_add over all 2^20
    points, edwards.rs:102-120 + 465-489; ~20-30 s on the box's host threads), not with the GPU's own scalar multiplications:
    plain call, both generator handles, and the 8-rank sharded decomposition through the in-process exchange kernel's
    sibling path (partials + fold)."""
    import os
    import torch
    from dusk_zerocaf_b200 import synth
    n = 1 << 20
    P = np.ascontiguousarray(big_points[:n])
    s = synth.synth_scalar(104, 0, n)
    want = oracle.msm_naive(P, s, threads=os.cpu_count() or THREADS)
    assert oracle.pt_is_valid(want)
    ctx = zc.default_context()
    L = ctx._L
    got = zc.batch.msm(P, s, window_bits=16)
    assert oracle.pt_eq(got, want)
    assert oracle.ris_compress(got) == oracle.ris_compress(want)
    dP = torch.from_numpy(P.view(np.int64)).cuda()
    dS = torch.from_numpy(s.view(np.int64)).cuda()
    out = torch.zeros(20, dtype=torch.int64, device="cuda")
    for kind in (zc.GEN_PREPARED, zc.GEN_FIXED_BASE):
        g = ctx.msm_generators(dP.data_ptr(), n, kind, 16, 0, 1)
        g.msm(dS.data_ptr(), out.data_ptr(), window_bits=16)
        ctx.sync()
        assert oracle.pt_eq(out.cpu().numpy().view(np.uint64), want), kind
        g.close()
    parts = torch.zeros((8, 20), dtype=torch.int64, device="cuda")
    g = ctx.msm_generators(dP.data_ptr(), n, zc.GEN_PREPARED)
    for r in range(8):
        g.msm_partial(dS.data_ptr(), parts[r].data_ptr(), r, 8, window_bits=16)
    ctx.check(L.zc_point_fold_dev(ctx._h, parts.data_ptr(), 8, out.data_ptr()))
    ctx.sync()
    assert oracle.pt_eq(out.cpu().numpy().view(np.uint64), want)
    g.close()
