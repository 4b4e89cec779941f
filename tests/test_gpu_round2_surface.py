"""Round-2 additions to the C ABI, each against the oracle / the reference's own vectors:
Div (field.rs:277-299, KAT :1242-1260), Scalar::into_bits (scalar.rs:352-366, KAT :979-1010), the reference's tests on
NON-canonical operands (field.rs:1160-1167 add_field_l, :1193-1200 subtract_field_l), opt-in input validation
(ZC_ERR_NONCANONICAL), and config 2 on the 32-byte wire format (field.rs:563-631)."""
import numpy as np
import pytest

from conftest import SEED, kat_arr

pytestmark = pytest.mark.gpu

P_INT = (1 << 252) + 27742317777372353535851937790883648493
L_INT = (1 << 249) + 14490550575682688738086195780655237219


@pytest.fixture(scope="module")
def zc():
    import dusk_zerocaf_b200 as z
    z.default_context()
    return z


def test_div_kat_and_random(zc, oracle, kats):
    b = zc.batch
    inl = kats["field"]["inline"]["division"]
    a = np.array([inl["a"], 0, 0, 0, 0], dtype=np.uint64)
    d = np.array([inl["b"], 0, 0, 0, 0], dtype=np.uint64)
    got = b.fe_div(b.fe_neg(a)[0], d)[0]                                   # -a / b, field.rs:1242-1260
    assert np.array_equal(got, np.array(inl["neg_a_over_b"], dtype=np.uint64))
    n = 3000
    x, y = oracle.synth_fe(SEED, 500, 0, n), oracle.synth_fe(SEED, 501, 0, n)
    y[0] = oracle.int_to_limbs(1)
    y[1] = oracle.int_to_limbs(P_INT - 1)
    want = np.stack([oracle.fe_div(x[i], y[i]) for i in range(n)])
    assert np.array_equal(b.fe_div(x, y), want)
    # x / y * y == x
    assert np.array_equal(b.fe_mul(b.fe_div(x, y), y), x)
    # the reference asserts on a zero divisor (field.rs:285); here the element is 0 and nothing aborts
    y[5] = 0
    assert not b.fe_div(x, y)[5].any()


def test_scalar_into_bits(zc, oracle, kats):
    b = zc.batch
    inl = kats["scalar"]["inline"]
    lm1 = oracle.int_to_limbs(L_INT - 1)
    assert list(b.scalar_into_bits(lm1)[0]) == inl["into_bits_minus_one"]    # scalar.rs:979-1010
    n = 2000
    s = oracle.synth_scalar(SEED, 510, 0, n)
    s[0] = 0
    s[1] = oracle.int_to_limbs(9)
    s[2] = oracle.int_to_limbs(1 << 249)
    got = b.scalar_into_bits(s)
    assert got.shape == (n, 256)
    for i in list(range(8)) + [n // 2, n - 1]:
        assert np.array_equal(got[i], oracle.sc_into_bits(s[i])), i
    # all of them against the value itself
    vals = [oracle.limbs_to_int(s[i]) for i in range(n)]
    packed = np.packbits(got, axis=1, bitorder="little")
    for i in range(n):
        assert int.from_bytes(packed[i].tobytes(), "little") == vals[i]


def test_reference_noncanonical_operand_cases(zc, oracle):
    """add_field_l / subtract_field_l (field.rs:1160-1167, 1193-1200): the reference feeds FIELD_L itself as an operand and
    expects 2 + p = 2 and 2 - p = 2.  The kernels are specified for canonical inputs, but these two cases hold too, and
    the oracle (a 1:1 restatement) agrees limb for limb."""
    b = zc.batch
    two = oracle.int_to_limbs(2)
    p = oracle.int_to_limbs(P_INT)
    assert np.array_equal(b.fe_add(two, p)[0], two)
    assert np.array_equal(b.fe_sub(two, p)[0], two)
    assert np.array_equal(oracle.fe_add_batch(two[None], p[None])[0], two)
    assert np.array_equal(oracle.fe_sub_batch(two[None], p[None])[0], two)


def test_check_canonical_and_validation_mode(zc, oracle):
    import torch
    b = zc.batch
    ctx = zc.default_context()
    n = 1000
    x, y = oracle.synth_fe(SEED, 520, 0, n), oracle.synth_fe(SEED, 521, 0, n)
    assert b.check_canonical("fe", x) is None
    bad = x.copy()
    bad[700] = oracle.int_to_limbs(P_INT)                                  # the modulus itself
    bad[900] = oracle.int_to_limbs(P_INT + 5)
    assert b.check_canonical("fe", bad) == 700
    bad2 = x.copy()
    bad2[321, 1] |= np.uint64(1 << 52)                                     # a limb with bit 52 set
    assert b.check_canonical("fe", bad2) == 321
    bad3 = x.copy()
    bad3[3, 4] = np.uint64(1 << 48)                                        # top limb beyond 48 bits (value >= 2^256)
    assert b.check_canonical("fe", bad3) == 3
    s = oracle.synth_scalar(SEED, 522, 0, n)
    assert b.check_canonical("scalar", s) is None
    sb = s.copy()
    sb[10] = oracle.int_to_limbs(L_INT)                                    # < p but >= L
    assert b.check_canonical("scalar", sb) == 10
    assert b.check_canonical("fe", sb) is None
    base = np.array(oracle.pt_scalar_mul_batch(np.tile(np.array(zc.synth.BASEPOINT), (64, 1)), s[:64], threads=4))
    assert b.check_canonical("point", base) is None
    pb = base.copy()
    pb[40, 12] |= np.uint64(1 << 60)                                       # inside Z of point 40
    assert b.check_canonical("point", pb) == 40
    # validation mode: the hot-path entry points refuse non-canonical inputs instead of computing garbage
    ctx.set_validation(True)
    try:
        assert np.array_equal(b.fe_mul(x, y), oracle.fe_mul_batch(x, y))   # canonical inputs: unchanged results
        assert np.array_equal(b.point_add(base, base[::-1].copy()), oracle.pt_add_batch(base, base[::-1].copy()))
        L = ctx._L
        out = np.empty_like(x)
        assert L.zc_fe_mul_batch(ctx._h, bad.ctypes.data, y.ctypes.data, out.ctypes.data, n) == 4
        assert "700" in L.zc_last_error_string(ctx._h).decode()
        assert L.zc_fe_add_batch(ctx._h, x.ctypes.data, bad2.ctypes.data, out.ctypes.data, n) == 4
        assert L.zc_scalar_mul_batch(ctx._h, sb.ctypes.data, s.ctypes.data, out.ctypes.data, n) == 4
        o2 = np.empty_like(x)
        assert L.zc_fe_mul_square_batch(ctx._h, x.ctypes.data, bad3.ctypes.data, out.ctypes.data, o2.ctypes.data, n) == 4
        po = np.empty_like(base)
        assert L.zc_point_add_batch(ctx._h, pb.ctypes.data, base.ctypes.data, po.ctypes.data, 64) == 4
        assert L.zc_point_scalar_mul_batch(ctx._h, base.ctypes.data, sb[:64].ctypes.data, po.ctypes.data, 64, 1) == 4
        # MSM: a scalar >= 2^250 would be masked into a wrong bucket without the check
        big = s[:64].copy()
        big[7] = oracle.int_to_limbs((1 << 251) + 12345)
        pt = np.empty(20, dtype=np.uint64)
        assert L.zc_msm(ctx._h, base.ctypes.data, big.ctypes.data, 64, 16, pt.ctypes.data) == 4
        assert L.zc_msm(ctx._h, base.ctypes.data, s[:64].ctypes.data, 64, 16, pt.ctypes.data) == 0
        assert oracle.pt_eq(pt, oracle.msm_naive(base, s[:64], threads=4))
        # _dev twins
        dx, dbad = torch.from_numpy(x.view(np.int64)).cuda(), torch.from_numpy(bad.view(np.int64)).cuda()
        dout = torch.empty_like(dx)
        assert L.zc_fe_mul_batch_dev(ctx._h, dx.data_ptr(), dbad.data_ptr(), dout.data_ptr(), n) == 4
        assert L.zc_fe_mul_batch_dev(ctx._h, dx.data_ptr(), dx.data_ptr(), dout.data_ptr(), n) == 0
    finally:
        ctx.set_validation(False)
    out = np.empty_like(x)
    assert ctx._L.zc_fe_mul_batch(ctx._h, bad.ctypes.data, y.ctypes.data, out.ctypes.data, n) == 0      # off again: no check


def test_mul_square_packed_wire_format(zc, oracle):
    b = zc.batch
    for n in (1, 255, 70_000):
        x, y = oracle.synth_fe(SEED, 530, 0, n), oracle.synth_fe(SEED, 531, 0, n)
        if n > 4:
            x[0] = 0
            x[1] = oracle.int_to_limbs(P_INT - 1)
            y[1] = oracle.int_to_limbs(P_INT - 1)
            x[2] = oracle.int_to_limbs(1)
        xb, yb = b.fe_to_bytes(x), b.fe_to_bytes(y)
        prod_b, sq_b = b.fe_mul_square_packed(xb, yb)
        prod, sq = oracle.fe_mul_batch(x, y), oracle.fe_square_batch(x)
        assert np.array_equal(b.fe_from_bytes(prod_b), prod), n
        assert np.array_equal(b.fe_from_bytes(sq_b), sq), n
