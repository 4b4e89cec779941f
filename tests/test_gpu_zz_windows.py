"""Every window size 8..16 against scalars that exercise the top windows: values in [2^249, L) put a carry into the window
that starts at or just below bit 250 (c = 10: a window starting exactly at bit 250 holds the digit 1), L - 1 and 2^249 - 1
give the extreme digits.  Plain Pippenger and the fixed-base tables, one GPU and a 3-rank sharding, against the oracle."""
import numpy as np
import pytest

from conftest import SEED

pytestmark = pytest.mark.gpu

L_INT = (1 << 249) + 14490550575682688738086195780655237219


@pytest.fixture(scope="module")
def zc():
    import dusk_zerocaf_b200 as z
    z.default_context()
    return z


def test_msm_all_window_sizes_with_top_bit_scalars(zc, oracle):
    import torch
    from test_gpu_parity import synth_points
    n = 200
    P = synth_points(oracle, 95, n)
    s = oracle.synth_scalar(SEED, 96, 0, n)
    edge = [L_INT - 1, L_INT - 2, 1 << 249, (1 << 249) + 12345, (1 << 249) - 1, (1 << 248), 0, 1, (1 << 249) + (1 << 123),
            ((1 << 249) - 1) ^ (1 << 240), (1 << 250) - (1 << 249) + 7]
    for j, v in enumerate(edge):
        s[j] = oracle.int_to_limbs(v % L_INT)
    for j in range(len(edge), 40):                                   # more top-bit-set scalars: 2^249 + small
        s[j] = oracle.int_to_limbs((1 << 249) + (int(oracle.limbs_to_int(s[j])) % (1 << 120)))
    want = oracle.msm_naive(P, s, threads=8)
    ctx = zc.default_context()
    Lb = ctx._L
    dP = torch.from_numpy(P.view(np.int64)).cuda()
    dS = torch.from_numpy(s.view(np.int64)).cuda()
    out = torch.zeros(20, dtype=torch.int64, device="cuda")
    for c in range(8, 17):
        ctx.check(Lb.zc_msm_dev(ctx._h, dP.data_ptr(), dS.data_ptr(), n, c, out.data_ptr()))
        ctx.sync()
        assert oracle.pt_eq(out.cpu().numpy().view(np.uint64), want), ("plain", c)
        for R in (1, 3):
            parts = torch.zeros((R, 20), dtype=torch.int64, device="cuda")
            for r in range(R):
                g = ctx.msm_generators(dP.data_ptr(), n, zc.GEN_FIXED_BASE, c, r, R)
                g.msm_partial(dS.data_ptr(), parts[r].data_ptr(), r, R)
                g.close()
            ctx.check(Lb.zc_point_fold_dev(ctx._h, parts.data_ptr(), R, out.data_ptr()))
            ctx.sync()
            assert oracle.pt_eq(out.cpu().numpy().view(np.uint64), want), ("fixed_base", c, R)
