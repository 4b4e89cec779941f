"""GPU parity tests proper: every call goes through the C ABI of libzerocaf_b200.so and is compared, bit for
bit, with the CPU oracle (oracle/) on the same seeded inputs and with the reference's own golden vectors
(tests/golden/reference_kats.json).  Point results of the fast paths / MSM are compared canonically."""
import numpy as np
import pytest

from conftest import SEED, kat_arr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def zc():
    import dusk_zerocaf_b200 as z
    z.default_context()
    return z


def F(kats, n): return kat_arr(kats, "field", n)
def S(kats, n): return kat_arr(kats, "scalar", n)
def E(kats, n): return kat_arr(kats, "edwards", n)
def C(kats, n): return kat_arr(kats, "constants", n)


def synth_points(oracle, stream, n, threads=8):
    """P_i = [r_i] B with r_i synthetic scalars; computed by the ORACLE (test input only)."""
    B = None
    import json, os
    with open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")) as f:
        k = json.load(f)
    B = np.array(k["constants"]["items"]["BASEPOINT"]["values"], dtype=np.uint64)
    r = oracle.synth_scalar(SEED, stream, 0, n)
    return oracle.pt_scalar_mul_batch(np.tile(B, (n, 1)), r, threads=threads)


# ---------------------------------------------------------------------------------------------------------
# field KATs through the ABI (reference field.rs:1136-1240, 1492-1522)
# ---------------------------------------------------------------------------------------------------------
def test_field_kats(zc, kats):
    b = zc.batch
    A, B, Cc = F(kats, "A"), F(kats, "B"), F(kats, "C")
    assert np.array_equal(b.fe_mul(A, B)[0], F(kats, "A_TIMES_B"))       # mul_with_modulo field.rs:1202
    assert np.array_equal(b.fe_mul(A, Cc)[0], F(kats, "A_TIMES_C"))      # mul_without_modulo :1210
    assert np.array_equal(b.fe_square(A)[0], F(kats, "A_SQUARE"))        # square :1218
    assert np.array_equal(b.fe_square(B)[0], F(kats, "B_SQUARE"))
    assert np.array_equal(b.fe_add(A, B)[0], F(kats, "A_PLUS_B"))        # addition_with_modulo :1136
    assert np.array_equal(b.fe_sub(A, B)[0], F(kats, "A_MINUS_B"))       # subtraction :1169
    assert np.array_equal(b.fe_sub(B, A)[0], F(kats, "B_MINUS_A"))
    assert np.array_equal(b.fe_neg(A)[0], F(kats, "MINUS_A"))            # neg :1492
    assert np.array_equal(b.fe_neg(B)[0], F(kats, "MINUS_B"))
    zero, one = np.zeros(5, np.uint64), np.array([1, 0, 0, 0, 0], np.uint64)
    minus_one = b.fe_neg(one)[0]
    assert np.array_equal(b.fe_add(minus_one, one)[0], zero)             # -1 + 1 = 0
    assert np.array_equal(b.fe_square(zero)[0], zero)                    # square_zero_and_identity :1231
    assert np.array_equal(b.fe_square(one)[0], one)
    assert np.array_equal(b.fe_neg(zero)[0], zero)


def test_scalar_kats(zc, kats):
    b = zc.batch
    X, Y = S(kats, "X"), S(kats, "Y")
    assert np.array_equal(b.scalar_mul(X, Y)[0], S(kats, "X_TIMES_Y"))   # scalar_mul scalar.rs:860
    assert np.array_equal(b.scalar_square(Y)[0], S(kats, "Y_SQ"))        # square :893
    A, B = S(kats, "A"), S(kats, "B")
    assert np.array_equal(b.scalar_sub(A, B)[0], S(kats, "AB"))          # sub :818
    assert np.array_equal(b.scalar_sub(B, A)[0], S(kats, "BA"))


def test_point_kats(zc, kats, oracle):
    b = zc.batch
    P1, P2 = E(kats, "P1_EXTENDED"), E(kats, "P2_EXTENDED")
    assert np.array_equal(b.point_add(P1, P2)[0], E(kats, "P4_EXTENDED"))   # extended_point_addition edwards.rs:1388 (limb-exact)
    dbl = b.point_double(P1)[0]
    assert oracle.pt_eq(dbl, E(kats, "P3_EXTENDED"))                        # extended_point_doubling :1394 (affine)
    assert np.array_equal(dbl, oracle.pt_double(P1))                        # limb-exact vs Double = self + self
    assert np.array_equal(b.point_neg(P1)[0], oracle.pt_neg(P1))
    assert np.array_equal(b.point_sub(P1, P2)[0], oracle.pt_sub(P1, P2))


# ---------------------------------------------------------------------------------------------------------
# BASELINE config 1: 1k random FieldElement::mul, bit-exact
# ---------------------------------------------------------------------------------------------------------
def test_cfg1_1k_field_mul(zc, oracle):
    a = oracle.synth_fe(SEED, 1, 0, 1000)
    bb = oracle.synth_fe(SEED, 2, 0, 1000)
    assert np.array_equal(zc.batch.fe_mul(a, bb), oracle.fe_mul_batch(a, bb))


@pytest.mark.parametrize("n", [1, 31, 257, 100_003])
def test_field_ops_random(zc, oracle, n):
    b = zc.batch
    a = oracle.synth_fe(SEED, 3, 0, n)
    c = oracle.synth_fe(SEED, 4, 0, n)
    # full-range operands too: a' = -a covers values up to p-1
    na = oracle.fe_neg_batch(a)
    for x, y in ((a, c), (na, c), (na, na)):
        assert np.array_equal(b.fe_mul(x, y), oracle.fe_mul_batch(x, y))
        assert np.array_equal(b.fe_add(x, y), oracle.fe_add_batch(x, y))
        assert np.array_equal(b.fe_sub(x, y), oracle.fe_sub_batch(x, y))
    assert np.array_equal(b.fe_square(na), oracle.fe_square_batch(na))
    assert np.array_equal(b.fe_neg(a), na)
    prod, sq = b.fe_mul_square(na, c)
    assert np.array_equal(prod, oracle.fe_mul_batch(na, c))
    assert np.array_equal(sq, oracle.fe_square_batch(na))


def test_field_edge_values(zc, oracle):
    p = 2**252 + 27742317777372353535851937790883648493
    vals = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, 2**252 - 1, 2**252, 2**251, 2**128 - 1, 2**52, 2**52 - 1]
    arr = np.array([oracle.int_to_limbs(v) for v in vals], dtype=np.uint64)
    n = len(vals)
    x = np.repeat(arr, n, axis=0)
    y = np.tile(arr, (n, 1))
    b = zc.batch
    assert np.array_equal(b.fe_mul(x, y), oracle.fe_mul_batch(x, y))
    assert np.array_equal(b.fe_add(x, y), oracle.fe_add_batch(x, y))
    assert np.array_equal(b.fe_sub(x, y), oracle.fe_sub_batch(x, y))
    assert np.array_equal(b.fe_square(arr), oracle.fe_square_batch(arr))


def test_scalar_ops_random(zc, oracle):
    b = zc.batch
    n = 5000
    a = oracle.synth_scalar(SEED, 5, 0, n)
    c = oracle.synth_scalar(SEED, 6, 0, n)
    na = oracle.sc_sub_batch(np.zeros_like(a), a)     # up to L-1
    for x, y in ((a, c), (na, c), (na, na)):
        assert np.array_equal(b.scalar_mul(x, y), oracle.sc_mul_batch(x, y))
        assert np.array_equal(b.scalar_add(x, y), oracle.sc_add_batch(x, y))
        assert np.array_equal(b.scalar_sub(x, y), oracle.sc_sub_batch(x, y))
    assert np.array_equal(b.scalar_square(na), oracle.sc_square_batch(na))
    assert np.array_equal(b.scalar_neg(a), na)


def test_empty_and_inplace(zc, oracle):
    b = zc.batch
    e = np.zeros((0, 5), np.uint64)
    assert b.fe_mul(e, e).shape == (0, 5)
    assert b.point_add(np.zeros((0, 20), np.uint64), np.zeros((0, 20), np.uint64)).shape == (0, 20)
    # in place: out aliases a
    a = oracle.synth_fe(SEED, 7, 0, 1000)
    c = oracle.synth_fe(SEED, 8, 0, 1000)
    want = oracle.fe_mul_batch(a, c)
    ctx = zc.default_context()
    ctx.call("zc_fe_mul_batch", a, c, a, 1000)
    assert np.array_equal(a, want)


# ---------------------------------------------------------------------------------------------------------
# config 3: point add / double, limb-exact
# ---------------------------------------------------------------------------------------------------------
def test_point_ops_random(zc, oracle):
    b = zc.batch
    n = 2048
    P = synth_points(oracle, 10, n)
    Q = synth_points(oracle, 11, n)
    assert np.array_equal(b.point_add(P, Q), oracle.pt_add_batch(P, Q, threads=8))
    assert np.array_equal(b.point_sub(P, Q), oracle.pt_sub_batch(P, Q, threads=8))
    assert np.array_equal(b.point_double(P), oracle.pt_double_batch(P, threads=8))
    assert np.array_equal(b.point_neg(P), oracle.pt_neg_batch(P))
    # identity and self-inverse edge cases
    I = np.tile(oracle.pt_identity(), (n, 1))
    assert np.array_equal(b.point_add(P, I), oracle.pt_add_batch(P, I))
    assert np.array_equal(b.point_add(I, I), oracle.pt_add_batch(I, I))
    assert np.array_equal(b.point_sub(P, P), oracle.pt_sub_batch(P, P))
    assert np.all(b.ristretto_eq(P, P) == 1)
    assert np.all(b.ristretto_eq(b.point_add(P, Q), b.point_add(Q, P)) == 1)
    assert not np.any(b.ristretto_eq(P, Q))


# ---------------------------------------------------------------------------------------------------------
# config 4: variable-base scalar-mul; strict = limb-exact vs double_and_add, fast = canonical
# ---------------------------------------------------------------------------------------------------------
def test_scalar_mul_strict(zc, oracle, kats):
    n = 300
    P = synth_points(oracle, 12, n)
    s = oracle.synth_scalar(SEED, 13, 0, n)
    # edge scalars: 0, 1, 2, L-1, 2^249-1
    L = 2**249 + 14490550575682688738086195780655237219
    for j, v in enumerate([0, 1, 2, L - 1, 2**249 - 1, 8, 2**248]):
        s[j] = oracle.int_to_limbs(v)
    got = zc.batch.point_scalar_mul(P, s, mode=0)
    want = oracle.pt_scalar_mul_batch(P, s, threads=8)
    assert np.array_equal(got, want)
    # unique_basepoint_test edwards.rs:1593: [L]B is the identity  (L as limbs is not a canonical scalar: use (L-1)B + B)
    B = C(kats, "BASEPOINT")
    lm1 = zc.batch.point_scalar_mul(B, oracle.int_to_limbs(L - 1), mode=0)[0]
    assert oracle.pt_eq(zc.batch.point_add(lm1, B)[0], oracle.pt_identity())


def test_scalar_mul_fast(zc, oracle):
    n = 300
    P = synth_points(oracle, 14, n)
    s = oracle.synth_scalar(SEED, 15, 0, n)
    L = 2**249 + 14490550575682688738086195780655237219
    for j, v in enumerate([0, 1, 2, L - 1, 2**249 - 1, 8, 2**248, 7, 9, 0x8888888888888888]):
        s[j] = oracle.int_to_limbs(v)
    got = zc.batch.point_scalar_mul(P, s, mode=1)
    want = oracle.pt_scalar_mul_batch(P, s, threads=8)
    for i in range(n):
        assert oracle.pt_eq(got[i], want[i]), i
        assert oracle.pt_is_valid(got[i])
    assert oracle.ris_compress(got[17]) == oracle.ris_compress(want[17])


def test_basepoint_mul_fixed_base(zc, oracle, kats):
    """[s]B through the fixed-base table == double_and_add(B, s) as a group element (edwards.rs:547-577): the reference's
    [0..15]B encodings, edge scalars ([L-1]B + B = O, unique_basepoint_test edwards.rs:1593-1600) and random scalars."""
    B = C(kats, "BASEPOINT")
    enc = kats["ristretto"]["small_multiples_hex"]
    L = 2**249 + 14490550575682688738086195780655237219
    edge = [0, 1, 2, 7, 8, 9, 15, 16, L - 1, L - 2, 2**249 - 1, 2**249, 2**248, 0x8888888888888888, 0x7777777777777777 << 180]
    s = np.concatenate([np.array([oracle.int_to_limbs(k) for k in list(range(16)) + edge], dtype=np.uint64),
                        oracle.synth_scalar(SEED, 76, 0, 400)])
    got = zc.batch.basepoint_mul(s)
    want = oracle.pt_scalar_mul_batch(np.tile(B, (s.shape[0], 1)), s, threads=8)
    for i in range(s.shape[0]):
        assert oracle.pt_is_valid(got[i]), i
        assert oracle.pt_eq(got[i], want[i]), i
    cg = zc.batch.ristretto_compress(got)
    if enc is not None:
        for k in range(16):
            assert cg[k].tobytes().hex() == enc[k], k
    assert oracle.pt_eq(oracle.pt_add(got[16 + 8], B), oracle.pt_identity())          # [L-1]B + B
    # agrees with the variable-base kernels on a larger batch
    s2 = oracle.synth_scalar(SEED, 77, 0, 5000)
    assert zc.batch.ristretto_eq(zc.batch.basepoint_mul(s2), zc.batch.point_scalar_mul(np.tile(B, (5000, 1)), s2, mode=1)).all()
    # more scalars than the persistent grid has threads (4 CTAs x 148 SMs x 128): every CTA loops, the last pass is ragged
    n3 = 4 * 148 * 128 * 2 + 12345
    s3 = oracle.synth_scalar(SEED, 78, 0, n3)
    got3 = zc.batch.basepoint_mul(s3)
    idx = np.unique(np.concatenate([np.arange(0, 40), np.arange(75776 - 20, 75776 + 20), np.arange(2 * 75776 - 20, 2 * 75776 + 20),
                                    np.arange(n3 - 40, n3), np.random.default_rng(5).integers(0, n3, 64)]))
    want3 = oracle.pt_scalar_mul_batch(np.tile(B, (idx.size, 1)), s3[idx], threads=8)
    for k, i in enumerate(idx):
        assert oracle.pt_eq(got3[i], want3[k]), int(i)
    assert zc.batch.ristretto_eq(got3, zc.batch.point_scalar_mul(np.tile(B, (n3, 1)), s3, mode=1)).all()


def test_ristretto_vectors_via_scalar_mul(zc, oracle, kats):
    """valid_encoding_test_vectors ristretto.rs:541-579: compress([k]B), k = 0..15."""
    B = C(kats, "BASEPOINT")
    enc = kats["ristretto"]["small_multiples_hex"]
    ks = np.array([oracle.int_to_limbs(k) for k in range(16)], dtype=np.uint64)
    for mode in (0, 1):
        got = zc.batch.point_scalar_mul(np.tile(B, (16, 1)), ks, mode=mode)
        acc = oracle.pt_identity()
        for k in range(16):
            assert oracle.ris_compress(got[k]) == oracle.ris_compress(acc), (mode, k)
            if enc is not None:
                assert oracle.ris_compress(got[k]).hex() == enc[k]
            acc = oracle.pt_add(acc, B)


# ---------------------------------------------------------------------------------------------------------
# batched canonicalisation (inverse, extended -> affine, Ristretto compress)
# ---------------------------------------------------------------------------------------------------------
def test_fe_invert(zc, oracle, kats):
    """savas_koc_inverse field.rs:1531-1547 (INV_MOD_A/B/C) + random values; a * a^-1 = 1."""
    b = zc.batch
    for nm in "ABC":
        assert np.array_equal(b.fe_invert(F(kats, nm))[0], F(kats, "INV_MOD_" + nm)), nm
    a = oracle.synth_fe(SEED, 71, 0, 512)
    a[0] = oracle.int_to_limbs(1)
    a[1] = b.fe_neg(oracle.int_to_limbs(1))[0]           # p - 1
    inv = b.fe_invert(a)
    one = np.tile(oracle.int_to_limbs(1), (512, 1))
    assert np.array_equal(b.fe_mul(a, inv), one)
    for i in range(0, 512, 37):
        assert np.array_equal(inv[i], oracle.fe_inverse(a[i])), i
    assert np.array_equal(b.fe_invert(np.zeros((1, 5), np.uint64))[0], np.zeros(5, np.uint64))


def test_point_to_affine(zc, oracle):
    """AffinePoint::from(EdwardsPoint) edwards.rs:1085-1092, limb-exact vs the oracle (Z != 1 inputs)."""
    P = synth_points(oracle, 72, 300)
    got = zc.batch.point_to_affine(P)
    want = oracle.pt_to_affine_batch(P, threads=8)
    assert np.array_equal(got, want)


def test_ristretto_compress(zc, oracle, kats):
    """RistrettoPoint::compress ristretto.rs:398-425: the reference's hex vectors for [0..15]B (ristretto.rs:541-579)
    and byte equality with the oracle on random points, their negations and torsion-shifted copies (P + P and -P)."""
    B = C(kats, "BASEPOINT")
    enc = kats["ristretto"]["small_multiples_hex"]
    acc = [oracle.pt_identity()]
    for k in range(1, 16):
        acc.append(oracle.pt_add(acc[-1], B))
    got = zc.batch.ristretto_compress(np.array(acc, dtype=np.uint64))
    for k in range(16):
        assert got[k].tobytes() == oracle.ris_compress(acc[k]), k
        if enc is not None:
            assert got[k].tobytes().hex() == enc[k], k
    P = synth_points(oracle, 73, 400)
    Q = np.concatenate([P, zc.batch.point_neg(P), zc.batch.point_double(P)])
    got = zc.batch.ristretto_compress(Q)
    want = oracle.ris_compress_batch(Q, threads=8)
    assert np.array_equal(got, want)
    # same group element, different representative (fast scalar-mul output) -> same bytes
    s = oracle.synth_scalar(SEED, 74, 0, 64)
    a = zc.batch.point_scalar_mul(P[:64], s, mode=0)
    f = zc.batch.point_scalar_mul(P[:64], s, mode=1)
    assert np.array_equal(zc.batch.ristretto_compress(a), zc.batch.ristretto_compress(f))


def test_ristretto_decompress_and_validity(zc, oracle, kats):
    """CompressedRistretto::decompress ristretto.rs:96-154 and ValidityCheck edwards.rs:393-400, limb-exact vs the oracle:
    the [0..15]B vectors (basepoint_compr_decompr :533, decompress_id :582), round trips of random points, and
    encodings the reference rejects (negative s, s >= p, random bytes: non-squares / negative t)."""
    b = zc.batch
    enc = kats["ristretto"]["small_multiples_hex"]
    rng = np.random.default_rng(20261017)
    P = synth_points(oracle, 75, 300)
    good = b.ristretto_compress(P)
    p = 2**252 + 27742317777372353535851937790883648493
    bad = []
    for k in range(40):
        v = int.from_bytes(good[k].tobytes(), "little")
        bad.append(np.frombuffer((p - v).to_bytes(32, "little"), dtype=np.uint8) if v else good[k])   # negative s
    bad.append(np.frombuffer((p + 5).to_bytes(32, "little"), dtype=np.uint8))                         # s >= p
    bad.append(np.full(32, 0xff, dtype=np.uint8))
    rnd = rng.integers(0, 256, size=(300, 32), dtype=np.uint8)
    rnd[:, 31] &= 0x0f                                                                               # many below (p-1)/2
    vec = [np.frombuffer(bytes.fromhex(h), dtype=np.uint8) for h in enc] if enc is not None else []
    allenc = np.concatenate([np.array(vec, dtype=np.uint8).reshape(-1, 32), good, np.array(bad, dtype=np.uint8), rnd])
    pts, ok = b.ristretto_decompress(allenc)
    n_some = 0
    for i in range(allenc.shape[0]):
        want = oracle.ris_decompress(allenc[i])
        assert (want is not None) == bool(ok[i]), i
        if want is not None:
            n_some += 1
            assert np.array_equal(pts[i], want), i
    assert n_some >= len(vec) + 300 and n_some < allenc.shape[0]          # both branches exercised
    # decode(encode(P)) is P up to 4-torsion: Ristretto-equal, and re-encodes to the same bytes
    dec = pts[len(vec):len(vec) + 300]
    assert b.ristretto_eq(dec, P).all()
    assert np.array_equal(b.ristretto_compress(dec), good)
    # validity: decoded points and extended-coordinate points satisfy the curve equation, a perturbed one does not
    assert b.point_is_valid(dec).all() and b.point_is_valid(P).all()
    Q = P[:8].copy(); Q[:, 0] ^= np.uint64(1)
    got = b.point_is_valid(Q)
    assert [int(x) for x in got] == [oracle.pt_is_valid(q) for q in Q] and not got.any()


def test_ristretto_elligator_and_from_uniform_bytes(zc, oracle, kats):
    """elligator_ristretto_flavor ristretto.rs:430-471 and from_uniform_bytes :493-507, all 20 limbs vs the oracle: the
    reference's Sage vector (elligator_vs_ristretto_sage :678-720), r0 = 0 / 1 / p-1, canonical and non-canonical inputs
    (from_bytes keeps 256 bits, field.rs:563-587), random 64-byte strings."""
    b = zc.batch
    r = kats["ristretto"]
    r0 = oracle.fe_from_bytes(bytes.fromhex(r["elligator_input_hex"]))
    p = 2**252 + 27742317777372353535851937790883648493
    rng = np.random.default_rng(7)
    ins = [r0, oracle.int_to_limbs(0), oracle.int_to_limbs(1), oracle.int_to_limbs(p - 1), oracle.int_to_limbs(2)]
    ins += list(oracle.synth_fe(SEED, 78, 0, 300))
    for _ in range(100):                                                    # values in [p, 2^256): limbs as from_bytes builds them
        v = int.from_bytes(rng.integers(0, 256, 32, dtype=np.uint8).tobytes(), "little") | (1 << 255)
        ins.append(np.array([(v >> (52 * i)) & ((1 << 52) - 1) for i in range(4)] + [v >> 208], dtype=np.uint64))
    ins = np.array(ins, dtype=np.uint64)
    got = b.ristretto_elligator(ins)
    for i in range(ins.shape[0]):
        assert np.array_equal(got[i], oracle.ris_elligator(ins[i])), i
    assert b.point_is_valid(got).all()
    want0 = oracle.ris_elligator(r0)
    assert oracle.ris_compress(got[0]) == oracle.ris_compress(want0)
    data = rng.integers(0, 256, size=(300, 64), dtype=np.uint8)
    data[0] = 0
    data[1] = 0xff
    pts = b.ristretto_from_uniform_bytes(data)
    for i in range(data.shape[0]):
        assert np.array_equal(pts[i], oracle.ris_from_uniform_bytes(data[i])), i
    assert b.point_is_valid(pts).all()


# ---------------------------------------------------------------------------------------------------------
# config 5: MSM (derived oracle: fold of double_and_add)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,c", [(1, 16), (2, 8), (33, 8), (1000, 10), (4096, 13), (4096, 16)])
def test_msm_vs_naive(zc, oracle, n, c):
    P = synth_points(oracle, 20 + c, n)
    s = oracle.synth_scalar(SEED, 40 + c, 0, n)
    got = zc.batch.msm(P, s, window_bits=c)
    want = oracle.msm_naive(P, s, threads=8)
    assert oracle.pt_is_valid(got)
    assert oracle.pt_eq(got, want)
    assert oracle.ris_compress(got) == oracle.ris_compress(want)


def test_msm_edges(zc, oracle, kats):
    B = C(kats, "BASEPOINT")
    L = 2**249 + 14490550575682688738086195780655237219
    # empty -> identity
    assert np.array_equal(zc.batch.msm(None, None), oracle.pt_identity())
    # all scalars zero -> identity element
    P = synth_points(oracle, 60, 64)
    z = np.zeros((64, 5), np.uint64)
    assert oracle.pt_eq(zc.batch.msm(P, z), oracle.pt_identity())
    # all scalars one -> chained Add
    one = np.tile(oracle.int_to_limbs(1), (64, 1))
    acc = oracle.pt_identity()
    for i in range(64):
        acc = oracle.pt_add(acc, P[i])
    assert oracle.pt_eq(zc.batch.msm(P, one), acc)
    # extreme scalars: L-1 everywhere; repeated points (all digits collide in one bucket)
    big = np.tile(oracle.int_to_limbs(L - 1), (64, 1))
    assert oracle.pt_eq(zc.batch.msm(P, big), oracle.msm_naive(P, big, threads=8))
    same = np.tile(B, (500, 1))
    s = np.tile(oracle.int_to_limbs(0x7fff_8000_ffff_0001_8000), (500, 1))
    assert oracle.pt_eq(zc.batch.msm(same, s, window_bits=16), oracle.msm_naive(same, s, threads=8))
    # [L-1]B + [1]B = identity
    pts = np.stack([B, B])
    sc = np.stack([oracle.int_to_limbs(L - 1), oracle.int_to_limbs(1)])
    assert oracle.pt_eq(zc.batch.msm(pts, sc), oracle.pt_identity())


@pytest.mark.parametrize("n", [31, 32, 33, 127, 129, 1025])
def test_msm_operand_pass_shapes(zc, oracle, n):
    """The operand pass converts 32 points per warp through a shared-memory tile (16-byte loads): ragged last tiles, and a
    point array that is only 8-byte aligned (falls back to the element-wise pass).  Arbitrary points, device pointers."""
    import torch
    P = synth_points(oracle, 90, n)
    s = oracle.synth_scalar(SEED, 91, 0, n)
    want = oracle.msm_naive(P, s, threads=8)
    ctx = zc.default_context()
    dS = torch.from_numpy(s.view(np.int64)).cuda()
    buf = torch.zeros(n * 20 + 2, dtype=torch.int64, device="cuda")       # cudaMalloc'd: 256-byte aligned
    out = torch.zeros(20, dtype=torch.int64, device="cuda")
    for off in (0, 1):                                                    # 0: 16-byte aligned, 1: 8 bytes off
        view = buf[off:off + n * 20]
        view.copy_(torch.from_numpy(P.view(np.int64).reshape(-1)))
        assert (view.data_ptr() % 16 == 0) == (off == 0)
        ctx.check(ctx._L.zc_msm_dev(ctx._h, view.data_ptr(), dS.data_ptr(), n, 16, out.data_ptr()))
        ctx.sync()
        assert oracle.pt_eq(out.cpu().numpy().view(np.uint64), want), (n, off)


def test_msm_sharding_partials_fold_to_full(zc, oracle):
    """The multi-GPU decomposition without NCCL: partial points of ranks 0..R-1 folded in rank order == full MSM."""
    import torch
    n, c = 3000, 16
    P = synth_points(oracle, 70, n)
    s = oracle.synth_scalar(SEED, 71, 0, n)
    want = oracle.msm_naive(P, s, threads=8)
    ctx = zc.default_context()
    dP = torch.from_numpy(P.view(np.int64)).cuda()
    dS = torch.from_numpy(s.view(np.int64)).cuda()
    for R in (1, 2, 4, 8, 16):
        parts = torch.zeros((R, 20), dtype=torch.int64, device="cuda")
        for r in range(R):
            ctx.check(ctx._L.zc_msm_partial_dev(ctx._h, dP.data_ptr(), dS.data_ptr(), n, c, r, R, parts[r].data_ptr()))
        out = torch.zeros(20, dtype=torch.int64, device="cuda")
        ctx.check(ctx._L.zc_point_fold_dev(ctx._h, parts.data_ptr(), R, out.data_ptr()))
        ctx.sync()
        got = out.cpu().numpy().view(np.uint64)
        assert oracle.pt_eq(got, want), R


# ---------------------------------------------------------------------------------------------------------
# host-side mirror types read like the reference's own tests
# ---------------------------------------------------------------------------------------------------------
def test_msm_call_sequence_reuses_workspace_and_graph(zc, oracle):
    """Back-to-back MSMs with changing sizes, windows, shard parameters and repeated identical calls: the grow-only
    workspace, the cached CUDA graph (replayed when the arguments repeat, re-recorded when they change) and the sharded
    partials must never leak state from one call into the next."""
    import torch
    ctx = zc.default_context()
    L = ctx._L
    nmax = 6000
    P = synth_points(oracle, 90, nmax)
    s = oracle.synth_scalar(SEED, 91, 0, nmax)
    dP = torch.from_numpy(P.view(np.int64)).cuda()
    dS = torch.from_numpy(s.view(np.int64)).cuda()
    out = torch.zeros(20, dtype=torch.int64, device="cuda")
    want = {}
    seq = [(6000, 16, 0, 1), (6000, 16, 0, 1), (777, 9, 0, 1), (6000, 16, 0, 1), (1, 16, 0, 1), (2049, 13, 1, 2),
           (2049, 13, 1, 2), (2049, 13, 0, 2), (6000, 8, 0, 1), (33, 16, 3, 4), (6000, 16, 0, 1)]
    for n, c, r, R in seq:
        ctx.check(L.zc_msm_partial_dev(ctx._h, dP.data_ptr(), dS.data_ptr(), n, c, r, R, out.data_ptr()))
        ctx.sync()
        got = out.cpu().numpy().view(np.uint64).copy()
        assert oracle.pt_is_valid(got), (n, c, r, R)
        if R == 1:
            if n not in want:
                want[n] = oracle.msm_naive(P[:n], s[:n], threads=8)
            assert oracle.pt_eq(got, want[n]), (n, c)
        else:                                                  # fold all ranks' partials for this (n, c, R)
            parts = torch.zeros((R, 20), dtype=torch.int64, device="cuda")
            for rr in range(R):
                ctx.check(L.zc_msm_partial_dev(ctx._h, dP.data_ptr(), dS.data_ptr(), n, c, rr, R, parts[rr].data_ptr()))
            ctx.sync()
            assert torch.equal(parts[r], out) or oracle.pt_eq(parts[r].cpu().numpy().view(np.uint64), got)
            tot = torch.zeros(20, dtype=torch.int64, device="cuda")
            ctx.check(L.zc_point_fold_dev(ctx._h, parts.data_ptr(), R, tot.data_ptr()))
            ctx.sync()
            if n not in want:
                want[n] = oracle.msm_naive(P[:n], s[:n], threads=8)
            assert oracle.pt_eq(tot.cpu().numpy().view(np.uint64), want[n]), (n, c, R)


def test_msm_prepared_points(zc, oracle):
    """zc_msm_generators (ZC_GEN_PREPARED): same result through the handle's cached operands, for several scalar vectors and
    window sizes, with other MSMs (plain, and through a second live handle) in between."""
    import torch
    n = 5000
    P = synth_points(oracle, 80, n)
    ctx = zc.default_context()
    L = ctx._L
    dP = torch.from_numpy(P.view(np.int64)).cuda()
    out = torch.zeros(20, dtype=torch.int64, device="cuda")
    gens = ctx.msm_generators(dP.data_ptr(), n, zc.GEN_PREPARED)
    assert gens.device_bytes == n * 128
    for k, c in enumerate((16, 12, 16, 9)):
        s = oracle.synth_scalar(SEED, 81 + k, 0, n)
        dS = torch.from_numpy(s.view(np.int64)).cuda()
        gens.msm(dS.data_ptr(), out.data_ptr(), window_bits=c)
        ctx.sync()
        assert oracle.pt_eq(out.cpu().numpy().view(np.uint64), oracle.msm_naive(P, s, threads=8)), (k, c)
    # a plain MSM over OTHER points and an MSM through a second handle in between
    P2 = synth_points(oracle, 85, 700)
    s2 = oracle.synth_scalar(SEED, 86, 0, 700)
    dP2, dS2 = torch.from_numpy(P2.view(np.int64)).cuda(), torch.from_numpy(s2.view(np.int64)).cuda()
    gens2 = ctx.msm_generators(dP2.data_ptr(), 700, zc.GEN_PREPARED)
    ctx.check(L.zc_msm_dev(ctx._h, dP2.data_ptr(), dS2.data_ptr(), 700, 16, out.data_ptr()))
    ctx.sync()
    want2 = oracle.msm_naive(P2, s2, threads=8)
    assert oracle.pt_eq(out.cpu().numpy().view(np.uint64), want2)
    gens2.msm(dS2.data_ptr(), out.data_ptr())
    ctx.sync()
    assert oracle.pt_eq(out.cpu().numpy().view(np.uint64), want2)
    gens.msm(dS.data_ptr(), out.data_ptr())
    ctx.sync()
    assert oracle.pt_eq(out.cpu().numpy().view(np.uint64), oracle.msm_naive(P, s, threads=8))
    # sharded partials through the handle fold to the same element
    parts = torch.zeros((4, 20), dtype=torch.int64, device="cuda")
    for r in range(4):
        gens.msm_partial(dS.data_ptr(), parts[r].data_ptr(), r, 4)
    ctx.check(L.zc_point_fold_dev(ctx._h, parts.data_ptr(), 4, out.data_ptr()))
    ctx.sync()
    assert oracle.pt_eq(out.cpu().numpy().view(np.uint64), oracle.msm_naive(P, s, threads=8))
    gens2.close()
    gens.close()
    ctx.check(L.zc_msm_dev(ctx._h, dP.data_ptr(), dS.data_ptr(), n, 16, out.data_ptr()))
    ctx.sync()
    assert oracle.pt_eq(out.cpu().numpy().view(np.uint64), oracle.msm_naive(P, s, threads=8))


def test_msm_generators_do_not_alias_a_recycled_address(zc, oracle):
    """ADVICE r1 / VERDICT r1 weak 3: prepared state used to be keyed by the raw device pointer, so freeing the points and
    getting the same address back with other generators silently returned the OLD result.  The handle owns a snapshot:
    overwriting the array in place (what a free + malloc at the same address amounts to) changes the plain call and not the
    handle; a new handle made from the same address sees the new points; stale handles after destroy do not match graphs."""
    import torch
    n = 2000
    Pa, Pb = synth_points(oracle, 90, n), synth_points(oracle, 91, n)
    s = oracle.synth_scalar(SEED, 92, 0, n)
    ctx = zc.default_context()
    L = ctx._L
    dP = torch.from_numpy(Pa.view(np.int64)).cuda()
    dS = torch.from_numpy(s.view(np.int64)).cuda()
    out = torch.zeros(20, dtype=torch.int64, device="cuda")
    want_a, want_b = oracle.msm_naive(Pa, s, threads=8), oracle.msm_naive(Pb, s, threads=8)
    for kind in (zc.GEN_PREPARED, zc.GEN_FIXED_BASE):
        dP.copy_(torch.from_numpy(Pa.view(np.int64)))
        g_old = ctx.msm_generators(dP.data_ptr(), n, kind, 16, 0, 1)
        g_old.msm(dS.data_ptr(), out.data_ptr())
        ctx.sync()
        assert oracle.pt_eq(out.cpu().numpy().view(np.uint64), want_a)
        dP.copy_(torch.from_numpy(Pb.view(np.int64)))           # same address, other generators
        torch.cuda.synchronize()
        ctx.check(L.zc_msm_dev(ctx._h, dP.data_ptr(), dS.data_ptr(), n, 16, out.data_ptr()))
        ctx.sync()
        assert oracle.pt_eq(out.cpu().numpy().view(np.uint64), want_b), "plain call must see the new points"
        g_old.msm(dS.data_ptr(), out.data_ptr())
        ctx.sync()
        assert oracle.pt_eq(out.cpu().numpy().view(np.uint64), want_a), "the handle owns its snapshot"
        g_old.close()
        g_new = ctx.msm_generators(dP.data_ptr(), n, kind, 16, 0, 1)
        g_new.msm(dS.data_ptr(), out.data_ptr())                 # a recorded graph of g_old must not be replayed
        ctx.sync()
        assert oracle.pt_eq(out.cpu().numpy().view(np.uint64), want_b)
        g_new.close()


def test_msm_fixed_base_tables(zc, oracle):
    """zc_msm_generators (ZC_GEN_FIXED_BASE): pre-scaled per-window tables, one merged bucket set, no doubling chain -- the same
    group element as the naive sum for several window sizes (aligned and not, short top windows), several scalar vectors
    (incl. the extreme digits of L - 1 and small scalars), sharded over R ranks, and unaffected by other MSMs between."""
    import torch
    n = 3000
    P = synth_points(oracle, 70, n)
    ctx = zc.default_context()
    L = ctx._L
    dP = torch.from_numpy(P.view(np.int64)).cuda()
    out = torch.zeros(20, dtype=torch.int64, device="cuda")
    svecs = [oracle.synth_scalar(SEED, 71, 0, n), oracle.synth_scalar(SEED, 72, 0, n)]
    svecs[1][0] = oracle.int_to_limbs((1 << 249) + 14490550575682688738086195780655237219 - 1)      # L - 1
    svecs[1][1] = 0
    svecs[1][2] = oracle.int_to_limbs(1)
    svecs[1][3] = oracle.int_to_limbs((1 << 249) - 1)
    want = [oracle.msm_naive(P, s, threads=8) for s in svecs]
    dSs = [torch.from_numpy(s.view(np.int64)).cuda() for s in svecs]
    for c, R in ((16, 1), (13, 1), (8, 1), (11, 1), (16, 2), (16, 8), (12, 3), (16, 32)):
        parts = [torch.zeros((R, 20), dtype=torch.int64, device="cuda") for _ in svecs]
        for r in range(R):
            g = ctx.msm_generators(dP.data_ptr(), n, zc.GEN_FIXED_BASE, c, r, R)
            for k, dS in enumerate(dSs):
                g.msm_partial(dS.data_ptr(), parts[k][r].data_ptr(), r, R)
                # replay of the recorded graph: the same group element (bucket order is up to the histogram atomics)
                g.msm_partial(dS.data_ptr(), out.data_ptr(), r, R)
                ctx.sync()
                assert oracle.pt_eq(out.cpu().numpy().view(np.uint64), parts[k][r].cpu().numpy().view(np.uint64)), (c, R, r)
            g.close()
        for k in range(len(svecs)):
            ctx.check(L.zc_point_fold_dev(ctx._h, parts[k].data_ptr(), R, out.data_ptr()))
            ctx.sync()
            assert oracle.pt_eq(out.cpu().numpy().view(np.uint64), want[k]), (c, R, k)
    # tables for (c=16, rank 0 of 1): another shape is refused (ZC_ERR_MODE), plain calls in between do not disturb them
    g = ctx.msm_generators(dP.data_ptr(), n, zc.GEN_FIXED_BASE, 16, 0, 1)
    dS = dSs[0]
    assert L.zc_msm_gen_dev(ctx._h, g._g, dS.data_ptr(), 12, out.data_ptr()) == 3
    assert L.zc_msm_gen_partial_dev(ctx._h, g._g, dS.data_ptr(), 16, 1, 2, out.data_ptr()) == 3
    for c, m in ((16, n), (12, n), (16, 700)):
        ctx.check(L.zc_msm_dev(ctx._h, dP.data_ptr(), dS.data_ptr(), m, c, out.data_ptr()))
        ctx.sync()
        w = want[0] if m == n else oracle.msm_naive(P[:m], svecs[0][:m], threads=8)
        assert oracle.pt_eq(out.cpu().numpy().view(np.uint64), w), (c, m)
        g.msm(dS.data_ptr(), out.data_ptr())
        ctx.sync()
        assert oracle.pt_eq(out.cpu().numpy().view(np.uint64), want[0]), (c, m)
    g.close()
    # tiny and ragged sizes through the merged bucket set
    for m, c, R in ((1, 16, 1), (2, 8, 1), (33, 16, 2), (257, 10, 1), (1000, 16, 8)):
        parts = torch.zeros((R, 20), dtype=torch.int64, device="cuda")
        for r in range(R):
            g = ctx.msm_generators(dP.data_ptr(), m, zc.GEN_FIXED_BASE, c, r, R)
            g.msm_partial(dS.data_ptr(), parts[r].data_ptr(), r, R)
            g.close()
        ctx.check(L.zc_point_fold_dev(ctx._h, parts.data_ptr(), R, out.data_ptr()))
        ctx.sync()
        assert oracle.pt_eq(out.cpu().numpy().view(np.uint64), oracle.msm_naive(P[:m], svecs[0][:m], threads=8)), (m, c, R)
    import ctypes
    h = ctypes.c_void_p()
    assert L.zc_msm_generators_create_dev(ctx._h, dP.data_ptr(), n, zc.GEN_FIXED_BASE, 17, 0, 1, ctypes.byref(h)) == 3
    assert L.zc_msm_generators_create_dev(ctx._h, dP.data_ptr(), n, zc.GEN_FIXED_BASE, 16, 2, 2, ctypes.byref(h)) == 2
    assert L.zc_msm_generators_create_dev(ctx._h, dP.data_ptr(), n, 7, 16, 0, 1, ctypes.byref(h)) == 3
    ctx.check(L.zc_msm_dev(ctx._h, dP.data_ptr(), dS.data_ptr(), n, 16, out.data_ptr()))
    ctx.sync()
    assert oracle.pt_eq(out.cpu().numpy().view(np.uint64), want[0])


def test_vector_ops_pow_half_bytes_naf(zc, oracle, kats):
    """Pow / Half / to_bytes / from_bytes / window NAF / sqrt_ratio_i through the ABI, bit-exact against
    the oracle on seeded inputs, edge values and the reference's KATs (field.rs:1262-1269, scalar.rs:913-934, 1023-1052)."""
    b = zc.batch
    u64 = lambda *v: np.array(v, dtype=np.uint64)
    P_INT = (1 << 252) + 27742317777372353535851937790883648493
    L_INT = (1 << 249) + 14490550575682688738086195780655237219
    lim = oracle.int_to_limbs
    # KATs
    assert np.array_equal(b.fe_pow(F(kats, "A"), F(kats, "C"))[0], F(kats, "A_POW_C"))
    assert np.array_equal(b.fe_pow(F(kats, "A"), F(kats, "B"))[0], F(kats, "A_POW_B"))
    assert np.array_equal(b.scalar_pow(S(kats, "A"), S(kats, "B"))[0], S(kats, "A_POW_B"))
    assert np.array_equal(b.scalar_half(S(kats, "Y"))[0], S(kats, "Y_HALF"))
    assert np.array_equal(b.scalar_half(S(kats, "A"))[0], u64(0, 0, 0, 1, 0))
    assert list(b.scalar_window_naf(u64(7, 0, 0, 0, 0), 2)[0][:4]) == [-1, 0, 0, 1]
    k = u64(1122334455, 0, 0, 0, 0)
    for w in range(2, 8):
        assert np.array_equal(b.scalar_window_naf(k, w)[0], oracle.sc_compute_window_naf(k, w)), w
    # seeded batches + edge values
    n = 257
    fa, fe_ = oracle.synth_fe(SEED, 60, 0, n), oracle.synth_fe(SEED, 61, 0, n)
    sa, se = oracle.synth_scalar(SEED, 62, 0, n), oracle.synth_scalar(SEED, 63, 0, n)
    for arr, m in ((fa, P_INT), (fe_, P_INT), (sa, L_INT), (se, L_INT)):
        arr[0] = lim(0); arr[1] = lim(1); arr[2] = lim(m - 1); arr[3] = lim(m - 2); arr[4] = lim(2)
    fe_[5] = lim(0); fa[5] = lim(0); se[5] = lim(0); sa[5] = lim(0)           # 0^0 = 1
    assert np.array_equal(b.fe_pow(fa, fe_), np.stack([oracle.fe_pow(fa[i], fe_[i]) for i in range(n)]))
    assert np.array_equal(b.scalar_pow(sa, se), np.stack([oracle.sc_pow(sa[i], se[i]) for i in range(n)]))
    assert np.array_equal(b.fe_half(fa), np.stack([oracle.fe_half(x) for x in fa]))
    assert np.array_equal(b.scalar_half(sa), np.stack([oracle.sc_half(x) for x in sa]))
    # wire format: to_bytes, from_bytes round trip, Scalar::from_bytes range check
    fb, sb = b.fe_to_bytes(fa), b.scalar_to_bytes(sa)
    assert np.array_equal(fb, np.stack([np.frombuffer(oracle.fe_to_bytes(x), dtype=np.uint8) for x in fa]))
    assert np.array_equal(sb, np.stack([np.frombuffer(oracle.sc_to_bytes(x), dtype=np.uint8) for x in sa]))
    assert np.array_equal(b.fe_from_bytes(fb), fa)
    back, ok = b.scalar_from_bytes(sb)
    assert np.array_equal(back, sa) and ok.all()
    rnd = np.random.default_rng(7).integers(0, 256, (64, 32), dtype=np.uint8)    # arbitrary 256-bit strings
    rnd[0] = np.frombuffer(int(L_INT).to_bytes(32, "little"), dtype=np.uint8)     # L itself: rejected
    rnd[1] = np.frombuffer(int(L_INT - 1).to_bytes(32, "little"), dtype=np.uint8)
    assert np.array_equal(b.fe_from_bytes(rnd), np.stack([oracle.fe_from_bytes(x) for x in rnd]))
    got, ok = b.scalar_from_bytes(rnd)
    for i in range(64):
        v = int.from_bytes(rnd[i].tobytes(), "little")
        assert ok[i] == (1 if v < L_INT else 0), i
        assert oracle.limbs_to_int(got[i]) == v, i
    # NAF on a batch incl. L - 1 and values next to L (the recoding's k + |d| wraps mod L there)
    sa[6] = lim(L_INT - 3); sa[7] = lim(L_INT - 5); sa[8] = lim((1 << 249) - 1)
    for w in (2, 3, 5, 7):
        assert np.array_equal(b.scalar_window_naf(sa, w), np.stack([oracle.sc_compute_window_naf(x, w) for x in sa])), w
    ctx = zc.default_context()
    out = np.empty((1, 256), dtype=np.int8)
    assert ctx._L.zc_scalar_window_naf_batch(ctx._h, sa.ctypes.data, 8, out.ctypes.data, 1) == 3
    assert ctx._L.zc_scalar_window_naf_batch(ctx._h, sa.ctypes.data, 1, out.ctypes.data, 1) == 3
    # sqrt_ratio_i (field.rs:443-491): squares, non-squares, u = 0, v = 0
    u, v = oracle.synth_fe(SEED, 64, 0, n), oracle.synth_fe(SEED, 65, 0, n)
    u[0] = lim(0); v[1] = lim(0); u[2] = lim(0); v[2] = lim(0); u[3] = lim(1); v[3] = lim(1); u[4] = lim(4); v[4] = lim(1)
    r, sq = b.fe_sqrt_ratio_i(u, v)
    for i in range(n):
        c, want = oracle.fe_sqrt_ratio_i(u[i], v[i])
        assert int(sq[i]) == int(c) and np.array_equal(r[i], want), i


def test_abi_error_convention(zc, oracle):
    """Status codes instead of panics (include/zerocaf_b200.h): argument errors > 0 with a message, n = 0 is a no-op,
    and a failed call leaves the context usable."""
    import torch
    ctx = zc.default_context()
    L = ctx._L
    a = torch.zeros((4, 5), dtype=torch.int64, device="cuda")
    out = torch.zeros((4, 20), dtype=torch.int64, device="cuda")
    assert L.zc_fe_mul_batch_dev(ctx._h, a.data_ptr(), None, a.data_ptr(), 4) == 1                      # ZC_ERR_NULL
    assert b"null" in L.zc_last_error_string(ctx._h)
    assert L.zc_fe_mul_batch_dev(ctx._h, None, None, None, 0) == 0                                      # empty batch
    assert L.zc_point_scalar_mul_batch_dev(ctx._h, out.data_ptr(), a.data_ptr(), out.data_ptr(), 4, 7) == 3   # ZC_ERR_MODE
    assert L.zc_msm_dev(ctx._h, out.data_ptr(), a.data_ptr(), 4, 17, out.data_ptr()) == 3               # window out of range
    assert L.zc_msm_dev(ctx._h, out.data_ptr(), a.data_ptr(), 4, 7, out.data_ptr()) == 3
    assert L.zc_fe_mul_batch_dev(ctx._h, a.data_ptr(), a.data_ptr(), a.data_ptr(), (1 << 31) + 1) == 2  # ZC_ERR_SIZE
    assert L.zc_msm_sharded_dev(ctx._h, out.data_ptr(), a.data_ptr(), 4, 16, out.data_ptr()) == 5       # no communicator
    with pytest.raises(zc.ZerocafError):
        ctx.check(L.zc_msm_dev(ctx._h, None, None, 4, 16, out.data_ptr()))
    # still usable
    one = np.tile(oracle.int_to_limbs(1), (3, 1))
    assert np.array_equal(zc.batch.fe_mul(one, one), one)


def test_operator_surface(zc, kats, oracle):
    A, B = zc.FieldElement(F(kats, "A")), zc.FieldElement(F(kats, "B"))
    assert (A * B) == zc.FieldElement(F(kats, "A_TIMES_B"))
    assert (A + B) == zc.FieldElement(F(kats, "A_PLUS_B"))
    assert (A - B) == zc.FieldElement(F(kats, "A_MINUS_B"))
    assert (-A) == zc.FieldElement(F(kats, "MINUS_A"))
    assert A.square() == zc.FieldElement(F(kats, "A_SQUARE"))
    assert A.pow(B) == zc.FieldElement(F(kats, "A_POW_B"))                  # a_pow_b field.rs:1262
    assert A.inverse() == zc.FieldElement(F(kats, "INV_MOD_A"))             # savas_koc_inverse :1531
    assert A.half() + A.half() == A
    ok, root = zc.FieldElement.sqrt_ratio_i(A.square(), zc.FieldElement.one())
    assert ok and root.square() == A.square()
    with pytest.raises(ZeroDivisionError):
        zc.FieldElement.zero().inverse()
    SA, SB, SY = zc.Scalar(S(kats, "A")), zc.Scalar(S(kats, "B")), zc.Scalar(S(kats, "Y"))
    assert SA.pow(SB) == zc.Scalar(S(kats, "A_POW_B"))                      # mod_pow scalar.rs:929
    assert SY.half() == zc.Scalar(S(kats, "Y_HALF"))                        # half :913
    assert list(zc.Scalar(oracle.int_to_limbs(7)).compute_NAF()[:4]) == [-1, 0, 0, 1]   # naf :1023
    P1, P2 = zc.EdwardsPoint(E(kats, "P1_EXTENDED")), zc.EdwardsPoint(E(kats, "P2_EXTENDED"))
    assert np.array_equal((P1 + P2).limbs, E(kats, "P4_EXTENDED"))
    assert P1.double() == zc.EdwardsPoint(E(kats, "P3_EXTENDED"))
    eight = zc.Scalar(oracle.int_to_limbs(8))
    assert P1 * eight == P1.double().double().double()                      # extended_double_and_add edwards.rs:1410
    assert eight * P1 == P1 * eight
    assert P1 + zc.EdwardsPoint.identity() == P1
    assert (P1 - P1) == zc.EdwardsPoint.identity()
    R1 = zc.RistrettoPoint(E(kats, "P1_EXTENDED"))
    assert R1 + R1 == R1.double() and not (R1 == R1.double())
    assert R1.compress() == oracle.ris_compress(E(kats, "P1_EXTENDED"))
    x, y = P1.to_affine()
    want = oracle.pt_to_affine(E(kats, "P1_EXTENDED"))
    assert np.array_equal(x.limbs, want[0:5]) and np.array_equal(y.limbs, want[5:10])
