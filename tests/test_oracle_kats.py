"""Pin the CPU oracle against every known-answer vector the reference's own unit tests hold for the
hot path.  Each test names the reference test it mirrors (file:line under
/root/reference).  Golden numbers come from tests/golden/reference_kats.json.
"""
import numpy as np
import pytest

from conftest import kat_arr
from oracle import pymodel as pm


def F(kats, name):
    return kat_arr(kats, "field", name)


def S(kats, name):
    return kat_arr(kats, "scalar", name)


def C(kats, name):
    return kat_arr(kats, "constants", name)


def E(kats, name):
    return kat_arr(kats, "edwards", name)


def u64(*v):
    return np.array(v, dtype=np.uint64)


ZERO, ONE, TWO = u64(0, 0, 0, 0, 0), u64(1, 0, 0, 0, 0), u64(2, 0, 0, 0, 0)


def eq(a, b):
    return np.array_equal(np.asarray(a, dtype=np.uint64), np.asarray(b, dtype=np.uint64))


# ------------------------------------------------------------------------------------------------
# constants (src/backend/u64/constants.rs) -- numeric sanity through the independent bigint model
# ------------------------------------------------------------------------------------------------
def test_constants_against_bigint(kats):
    assert pm.from_limbs(C(kats, "FIELD_L")) == pm.P
    assert pm.from_limbs(C(kats, "L")) == pm.L
    assert pm.from_limbs(C(kats, "EDWARDS_D")) == pm.D
    assert pm.from_limbs(C(kats, "EDWARDS_A")) == pm.P - 1
    assert pm.from_limbs(C(kats, "RR_FIELD")) == pow(2, 520, pm.P)
    assert pm.from_limbs(C(kats, "RR")) == pow(2, 520, pm.L)
    assert (int(C(kats, "LFACTOR_FIELD")[0]) * pm.P) % 2**52 == 2**52 - 1
    assert (int(C(kats, "LFACTOR")[0]) * pm.L) % 2**52 == 2**52 - 1
    assert pow(pm.from_limbs(C(kats, "SQRT_MINUS_ONE")), 2, pm.P) == pm.P - 1
    i = pm.from_limbs(C(kats, "INV_SQRT_A_MINUS_D"))
    assert i * i % pm.P * ((-1 - pm.D) % pm.P) % pm.P == 1
    assert pm.from_limbs(C(kats, "POS_RANGE")) == (pm.P - 1) // 2
    assert pm.from_limbs(C(kats, "INVERSE_MOD_TWO")) == pow(2, -1, pm.P)
    assert pm.from_limbs(C(kats, "SCALAR_INVERSE_MOD_TWO")) == pow(2, -1, pm.L)
    B = pm.pt_from_limbs(C(kats, "BASEPOINT"))
    assert pm.on_curve(B) and B[1] == 3 * pow(5, -1, pm.P) % pm.P and B[0] * B[1] % pm.P == B[3]


# ------------------------------------------------------------------------------------------------
# FieldElement (src/backend/u64/field.rs tests :1136-1556)
# ------------------------------------------------------------------------------------------------
def test_field_addition(oracle, kats):
    minus_one = pm.to_limbs(pm.P - 1)
    assert eq(oracle.fe_add(minus_one, ONE), ZERO)                         # addition_with_modulo :1136
    assert eq(oracle.fe_add(F(kats, "A"), F(kats, "B")), F(kats, "A_PLUS_B"))  # addition_without_modulo :1144
    assert eq(oracle.fe_add(TWO, C(kats, "FIELD_L")), TWO)                 # add_field_l :1160


def test_field_subtraction(oracle, kats):
    A, B = F(kats, "A"), F(kats, "B")
    assert eq(oracle.fe_sub(A, B), F(kats, "A_MINUS_B"))                    # subtraction_with_mod :1169
    assert eq(oracle.fe_sub(B, A), F(kats, "B_MINUS_A"))                    # subtraction_without_mod :1177
    assert eq(oracle.fe_sub(B, B), ZERO)                                    # subtract_equals :1185
    assert eq(oracle.fe_sub(TWO, C(kats, "FIELD_L")), TWO)                  # subtract_field_l :1193


def test_field_mul_square(oracle, kats):
    A, B, Cc = F(kats, "A"), F(kats, "B"), F(kats, "C")
    assert eq(oracle.fe_mul(A, B), F(kats, "A_TIMES_B"))                    # mul_with_modulo :1202
    assert eq(oracle.fe_mul(A, Cc), F(kats, "A_TIMES_C"))                   # mul_without_modulo :1210
    assert eq(oracle.fe_square(A), F(kats, "A_SQUARE"))                     # square :1218
    assert eq(oracle.fe_square(B), F(kats, "B_SQUARE"))
    assert eq(oracle.fe_square(ZERO), ZERO) and eq(oracle.fe_square(ONE), ONE)  # square_zero_and_identity :1231


def test_field_division_pow(oracle, kats):
    d = kats["field"]["inline"]["division"]
    a, b = u64(d["a"], 0, 0, 0, 0), u64(d["b"], 0, 0, 0, 0)
    assert eq(oracle.fe_div(oracle.fe_neg(a), b), d["neg_a_over_b"])        # division :1242
    with pytest.raises(ZeroDivisionError):
        oracle.fe_div(a, ZERO)
    assert eq(oracle.fe_pow(F(kats, "A"), F(kats, "C")), F(kats, "A_POW_C"))  # a_pow_b :1262
    assert eq(oracle.fe_pow(F(kats, "A"), F(kats, "B")), F(kats, "A_POW_B"))


def test_field_sqrt_stack(oracle, kats):
    assert oracle.fe_legendre_symbol(F(kats, "A")) == 0                     # legendre_symbol :1271
    assert oracle.fe_legendre_symbol(u64(17, 0, 0, 0, 0)) == 1
    inp = u64(17, 0, 0, 0, 0)
    assert eq(oracle.fe_mod_sqrt(inp, 0), F(kats, "SQRT1_27_NEG"))          # mod_sqrt_tonelli_shanks :1281
    assert eq(oracle.fe_mod_sqrt(inp, 1), F(kats, "SQRT1_27_POS"))
    assert eq(oracle.fe_mod_sqrt(ZERO, 0), ZERO) and eq(oracle.fe_mod_sqrt(ZERO, 1), ZERO)
    _, r = oracle.fe_inv_sqrt(u64(27, 0, 0, 0, 0))                          # inv_sqrt :1298
    assert eq(oracle.fe_neg(r), F(kats, "INV_SQRT_27"))
    assert oracle.fe_mod_sqrt(F(kats, "A"), 0) is None                      # non_QRmod_sqrt :1306
    assert oracle.fe_mod_sqrt(F(kats, "A"), 1) is None


def test_field_bytes(oracle, kats):
    mob = bytes(int(x) for x in F(kats, "MINUS_ONE_BYTES"))
    minus_one = pm.to_limbs(pm.P - 1)
    assert eq(oracle.fe_from_bytes(mob), minus_one)                         # from_bytes_conversion :1313
    assert oracle.fe_to_bytes(minus_one) == mob                             # to_bytes_conversion :1321
    v = kats["field"]["inline"]["from_bytes_vector"]                        # from_ristretto255scalar :1378
    assert eq(oracle.fe_from_bytes(bytes(v["bytes"])), v["limbs"])
    assert oracle.fe_to_bytes(v["limbs"]) == bytes(v["bytes"])              # into_ristretto255scalar :1401
    assert oracle.fe_to_bytes(C(kats, "FIELD_L"))[31] < 0x80                # l_field_high_bit :1524


def test_field_two_pow_k_ord_half(oracle, kats):
    assert eq(oracle.fe_two_pow_k(0), ONE)                                  # two_pow_k :1424
    assert eq(oracle.fe_two_pow_k(252), F(kats, "TWO_POW_252"))
    assert eq(oracle.fe_two_pow_k(197), F(kats, "TWO_POW_197"))
    assert eq(oracle.fe_two_pow_k(104), F(kats, "TWO_POW_104"))
    assert oracle.fe_cmp(TWO, u64(0, 2, 0, 0, 0)) < 0                       # ord_impl :1450
    assert oracle.fe_cmp(u64(0, 0, 0, 0, 1), u64(0, 2498436546, 6587652167965486, 0, 0)) > 0
    assert oracle.fe_cmp(u64(0, 1, 2, 3, 4), u64(0, 1, 2, 3, 4)) == 0
    assert eq(oracle.fe_half_without_mod(u64(0, 1, 0, 0, 0)), u64(2251799813685248, 0, 0, 0, 0))  # half :1460
    assert eq(oracle.fe_half_without_mod(F(kats, "A_MINUS_B")), F(kats, "A_MINUS_B_HALF"))


def test_field_montgomery_neg_inverse(oracle, kats):
    A, B, Cc = F(kats, "A"), F(kats, "B"), F(kats, "C")
    assert eq(oracle.fe_to_montgomery(A), F(kats, "INV_MONT_A"))            # to_montgomery_conv :1476
    assert eq(oracle.fe_from_montgomery(F(kats, "INV_MONT_A")), A)          # from_montgomery_conv :1484
    assert eq(oracle.fe_neg(A), F(kats, "MINUS_A"))                         # negation :1492
    assert eq(oracle.fe_neg(B), F(kats, "MINUS_B"))
    minus_one = pm.to_limbs(pm.P - 1)
    assert eq(oracle.fe_neg(ONE), minus_one) and eq(oracle.fe_neg(minus_one), ONE)  # negate_one :1502
    assert eq(oracle.fe_neg(ZERO), ZERO)                                    # negate_zero :1516
    assert eq(oracle.fe_inverse(A), F(kats, "INV_MOD_A"))                   # savas_koc_inverse :1531
    assert eq(oracle.fe_inverse(B), F(kats, "INV_MOD_B"))
    assert eq(oracle.fe_inverse(Cc), F(kats, "INV_MOD_C"))
    with pytest.raises(ZeroDivisionError):
        oracle.fe_inverse(ZERO)


def test_field_random_vs_bigint(oracle):
    """20 000-sample cross-check of mul/square/add/sub/neg against `% p` (second opinion)."""
    n = 20000
    a = oracle.synth_fe(1, 0, 0, n)
    b = oracle.synth_fe(1, 1, 0, n)
    prod = oracle.fe_mul_batch(a, b)
    sq = oracle.fe_square_batch(a)
    add = oracle.fe_add_batch(a, b)
    sub = oracle.fe_sub_batch(a, b)
    neg = oracle.fe_neg_batch(a)
    for i in range(n):
        x, y = pm.from_limbs(a[i]), pm.from_limbs(b[i])
        assert x < 2**251 and y < 2**251
        assert pm.from_limbs(prod[i]) == x * y % pm.P
        assert pm.from_limbs(sq[i]) == x * x % pm.P
        assert pm.from_limbs(add[i]) == (x + y) % pm.P
        assert pm.from_limbs(sub[i]) == (x - y) % pm.P
        assert pm.from_limbs(neg[i]) == (-x) % pm.P
    assert int(prod.max()) < 2**52


# ------------------------------------------------------------------------------------------------
# Scalar (src/backend/u64/scalar.rs tests :786-1052)
# ------------------------------------------------------------------------------------------------
def test_scalar_add_sub(oracle, kats):
    A, B, AB, BA = S(kats, "A"), S(kats, "B"), S(kats, "AB"), S(kats, "BA")
    assert eq(oracle.sc_add(AB, BA), ZERO)                                  # add_with_modulo :800
    assert eq(oracle.sc_add(BA, A), B)                                      # add_without_modulo :810
    assert eq(oracle.sc_sub(A, B), AB)                                      # sub_with_modulo :819
    assert eq(oracle.sc_sub(B, A), BA)                                      # sub_without_modulo :827


def test_scalar_mul_square_montgomery(oracle, kats):
    X, Y = S(kats, "X"), S(kats, "Y")
    assert eq(oracle.sc_to_montgomery(S(kats, "A")), S(kats, "A_MONT"))     # to_montgomery_conversion :844
    assert eq(oracle.sc_from_montgomery(S(kats, "Y_MONT")), Y)              # from_montgomery_conversion :852
    assert eq(oracle.sc_mul(X, Y), S(kats, "X_TIMES_Y"))                    # scalar_mul :860
    assert eq(oracle.sc_mul(Y, ONE), Y)                                     # mul_by_identity :868
    assert eq(oracle.sc_mul(Y, ZERO), ZERO)                                 # mul_by_zero :877
    assert eq(oracle.sc_montgomery_mul(X, Y), S(kats, "X_TIMES_Y_MONT"))    # montgomery_mul :885
    assert eq(oracle.sc_square(Y), S(kats, "Y_SQ"))                         # square :893
    assert eq(oracle.sc_square(ZERO), ZERO) and eq(oracle.sc_square(ONE), ONE)  # :902


def test_scalar_half_pow_shr(oracle, kats):
    A = S(kats, "A")
    assert eq(oracle.sc_half(S(kats, "Y")), S(kats, "Y_HALF"))              # half :913
    assert eq(oracle.sc_half(A), u64(0, 0, 0, 1, 0))
    assert eq(oracle.sc_half(oracle.sc_half(A)), u64(0, 0, 2251799813685248, 0, 0))
    assert eq(oracle.sc_pow(A, S(kats, "B")), S(kats, "A_POW_B"))           # mod_pow :929
    assert eq(oracle.sc_two_pow_k(0), ONE)                                  # two_pow_k :951
    assert eq(oracle.sc_two_pow_k(249), u64(0, 0, 0, 0, 2199023255552))
    assert eq(oracle.sc_two_pow_k(248), u64(0, 0, 0, 0, 1099511627776))
    assert eq(oracle.sc_shr(A, 1), u64(0, 0, 0, 1, 0))                      # shr :962
    assert eq(oracle.sc_shr(u64(0, 0, 0, 1, 0), 1), u64(0, 0, 2251799813685248, 0, 0))
    assert eq(oracle.sc_shr(ONE, 1), ZERO) and eq(oracle.sc_shr(ZERO, 1), ZERO)
    minus_one = pm.to_limbs(pm.L - 1)
    assert eq(oracle.sc_shr(minus_one, 250), ZERO)
    assert eq(oracle.sc_shr(oracle.sc_two_pow_k(249), 248), TWO)
    assert eq(oracle.sc_shr(oracle.sc_two_pow_k(249), 249), ONE)


def test_scalar_bits_naf(oracle, kats):
    inl = kats["scalar"]["inline"]
    bits = oracle.sc_into_bits(pm.to_limbs(pm.L - 1))                       # into_bits :979
    assert list(bits) == inl["into_bits_minus_one"]
    assert not oracle.sc_into_bits(ZERO).any()
    nine = oracle.sc_into_bits(u64(9, 0, 0, 0, 0))
    assert nine[0] == 1 and nine[3] == 1 and nine.sum() == 2
    b249 = oracle.sc_into_bits(oracle.sc_two_pow_k(249))
    assert b249[249] == 1 and b249.sum() == 1
    assert list(oracle.sc_compute_naf(u64(7, 0, 0, 0, 0))[:4]) == [-1, 0, 0, 1]  # naf :1023
    k = u64(1122334455, 0, 0, 0, 0)                                         # window_naf :1029
    for w in (2, 3, 4, 5, 6):
        exp = inl["wnaf%d_1122334455" % w]
        assert list(oracle.sc_compute_window_naf(k, w)[:len(exp)]) == exp, w


def test_scalar_bytes(oracle):
    minus_one = pm.to_limbs(pm.L - 1)
    by = oracle.sc_to_bytes(minus_one)
    assert int.from_bytes(by, "little") == pm.L - 1
    assert eq(oracle.sc_from_bytes(by), minus_one)
    with pytest.raises(ValueError):                                         # assert :465
        oracle.sc_from_bytes(pm.L.to_bytes(32, "little"))


def test_scalar_random_vs_bigint(oracle):
    n = 5000
    a = oracle.synth_scalar(2, 0, 0, n)
    b = oracle.synth_scalar(2, 1, 0, n)
    prod = oracle.sc_mul_batch(a, b)
    sq = oracle.sc_square_batch(a)
    add = oracle.sc_add_batch(a, b)
    sub = oracle.sc_sub_batch(a, b)
    for i in range(n):
        x, y = pm.from_limbs(a[i]), pm.from_limbs(b[i])
        assert x < 2**249 and y < 2**249
        assert pm.from_limbs(prod[i]) == x * y % pm.L
        assert pm.from_limbs(sq[i]) == x * x % pm.L
        assert pm.from_limbs(add[i]) == (x + y) % pm.L
        assert pm.from_limbs(sub[i]) == (x - y) % pm.L


# ------------------------------------------------------------------------------------------------
# EdwardsPoint (src/edwards.rs tests :1354-1617)
# ------------------------------------------------------------------------------------------------
def test_point_neg_identity(oracle):
    ident = oracle.pt_identity()
    assert eq(ident, [0] * 5 + [1, 0, 0, 0, 0] + [1, 0, 0, 0, 0] + [0] * 5)
    assert oracle.pt_eq(oracle.pt_neg(ident), ident) == 1                   # extended_point_neg :1367


def test_point_addition_limb_exact(oracle, kats):
    P1, P2, P4 = E(kats, "P1_EXTENDED"), E(kats, "P2_EXTENDED"), E(kats, "P4_EXTENDED")
    res = oracle.pt_add(P1, P2)                                             # extended_point_addition :1388
    assert eq(res, P4)                                                      # limb-exact
    assert oracle.pt_eq(res, P4) == 1
    assert eq(pm.pt_to_limbs(pm.pt_add(pm.pt_from_limbs(P1), pm.pt_from_limbs(P2))), P4)


def test_point_doubling_affine(oracle, kats):
    P1, P3 = E(kats, "P1_EXTENDED"), E(kats, "P3_EXTENDED")
    assert oracle.pt_eq(oracle.pt_add(P1, P1), P3) == 1                     # doubling_by_addition :1394
    assert oracle.pt_eq(oracle.pt_double(P1), P3) == 1                      # extended_point_doubling :1400
    ident = oracle.pt_identity()
    assert oracle.pt_eq(oracle.pt_double(ident), ident) == 1
    eight = u64(8, 0, 0, 0, 0)                                              # extended_double_and_add :1410
    expect = oracle.pt_double(oracle.pt_double(oracle.pt_double(P1)))
    assert oracle.pt_eq(oracle.pt_double_and_add(P1, eight), expect) == 1


def test_point_generation_compression(oracle, kats):
    P1, P2 = E(kats, "P1_EXTENDED"), E(kats, "P2_EXTENDED")
    assert oracle.pt_eq(oracle.pt_new_from_y_coord(P2[5:10], 0), P2) == 1   # extended_point_generation :1420
    assert oracle.pt_eq(oracle.pt_new_from_y_coord(P1[5:10], 0), P1) == 1
    assert oracle.pt_new_from_y_coord(u64(15, 0, 0, 0, 0), 0) is None
    inl = kats["edwards"]["inline"]
    assert oracle.pt_compress(P1) == bytes(inl["P1_compress"])              # point_compression :1549
    assert oracle.pt_compress(P2) == bytes(inl["P2_compress"])
    c1 = bytes(int(x) for x in E(kats, "P1_COMPRESSED"))
    c2 = bytes(int(x) for x in E(kats, "P2_COMPRESSED"))
    assert oracle.pt_eq(oracle.pt_decompress(c1), P1) == 1                  # point_decompression :1564
    assert oracle.pt_eq(oracle.pt_decompress(c2), P2) == 1
    bad = bytes([250, 144, 188, 47, 13, 101, 118, 114, 201, 185, 169, 115, 255, 111, 40, 25, 69, 105,
                 170, 255, 113, 65, 120, 126, 170, 192, 48, 109, 112, 20, 221, 149])
    assert oracle.pt_decompress(bad) is None


def test_point_validity(oracle, kats):
    for name in ("P1_EXTENDED", "P2_EXTENDED", "P4_EXTENDED"):              # validity_check :1579
        assert oracle.pt_is_valid(E(kats, name)) == 1
    assert oracle.pt_is_valid(oracle.pt_identity()) == 1
    bad = E(kats, "P1_EXTENDED").copy()
    bad[0] += 1
    assert oracle.pt_is_valid(bad) == 0


def test_unique_basepoint(oracle, kats):
    y = oracle.fe_div(u64(3, 0, 0, 0, 0), u64(5, 0, 0, 0, 0))               # unique_basepoint_test :1593
    basep = oracle.pt_new_from_y_coord(y, 0)
    assert oracle.pt_is_valid(basep) == 1
    assert oracle.pt_eq(oracle.pt_double_and_add(basep, C(kats, "L")), oracle.pt_identity()) == 1
    assert oracle.pt_eq(basep, C(kats, "BASEPOINT")) == 1


def test_scalar_mul_algorithms_agree(oracle, kats):
    P1 = E(kats, "P1_EXTENDED")
    s215 = oracle.sc_two_pow_k(215)                                         # left_to_right_bin_mul :1602
    assert oracle.pt_eq(oracle.pt_double_and_add(P1, s215), oracle.pt_ltr_bin_mul(P1, s215)) == 1
    for s in (oracle.sc_two_pow_k(7), s215,                                 # naf_bin_mul :1607
              oracle.sc_sub(oracle.sc_two_pow_k(249), ONE), pm.to_limbs(pm.L - 1)):
        assert oracle.pt_eq(oracle.pt_double_and_add(P1, s), oracle.pt_binary_naf_mul(P1, s)) == 1


def test_double_and_add_limb_exact_vs_bigint(oracle, kats):
    B = C(kats, "BASEPOINT")
    sc = oracle.synth_scalar(3, 0, 0, 8)
    for i in range(8):
        got = oracle.pt_double_and_add(B, sc[i])
        want = pm.pt_double_and_add(pm.pt_from_limbs(B), pm.from_limbs(sc[i]))
        assert eq(got, pm.pt_to_limbs(want))


def test_odd_multiples_table(oracle, kats):
    """BASEPOINT_ODD_MULTIPLES_TABLE[i] == (2i-1)B for i >= 1, [0] == identity (constants.rs:216-972)."""
    tab = C(kats, "BASEPOINT_ODD_MULTIPLES_TABLE").reshape(126, 20)
    B = C(kats, "BASEPOINT")
    assert oracle.pt_eq(tab[0], oracle.pt_identity()) == 1
    for i in (1, 2, 3, 11, 64, 125):
        k = u64(2 * i - 1, 0, 0, 0, 0)
        assert oracle.pt_eq(oracle.pt_double_and_add(B, k), tab[i]) == 1, i


# ------------------------------------------------------------------------------------------------
# RistrettoPoint (src/ristretto.rs tests :526-721)
# ------------------------------------------------------------------------------------------------
def test_ristretto_small_multiples(oracle, kats):
    B = C(kats, "BASEPOINT")
    P = oracle.pt_identity()                                                # valid_encoding_test_vectors :541
    for k, hexenc in enumerate(kats["ristretto"]["small_multiples_hex"]):
        assert oracle.ris_compress(P).hex() == hexenc, k
        via_mul = oracle.pt_double_and_add(B, u64(k, 0, 0, 0, 0))
        assert oracle.ris_compress(via_mul).hex() == hexenc, k
        assert oracle.ris_eq(via_mul, P) == 1
        P = oracle.pt_add(P, B)


def test_ristretto_basepoint_roundtrip(oracle, kats):
    B = C(kats, "BASEPOINT")
    comp = oracle.ris_compress(B)                                           # basepoint_compr_decompr :533
    assert comp == bytes(kats["ristretto"]["compressed_basepoints"]["RISTRETTO_BASEPOINT_COMPRESSED"])
    dec = oracle.ris_decompress(comp)
    assert dec is not None and oracle.ris_eq(dec, B) == 1
    coset = C(kats, "FOUR_COSET_GROUP").reshape(4, 20)                      # four_coset_eq_basepoint :632
    for j in range(3):
        assert oracle.ris_eq(oracle.pt_add(B, coset[j]), B) == 1
    # four_torsion_diff :597: B - decompress(compress(B)) has order dividing 4
    diff = oracle.pt_sub(B, dec)
    four = oracle.pt_double_and_add(diff, u64(4, 0, 0, 0, 0))
    assert oracle.pt_compress(four) == bytes([1] + [0] * 31)


def test_ristretto_elligator(oracle, kats):
    r = kats["ristretto"]                                                   # elligator_vs_ristretto_sage :678
    expected = np.array(r["elligator_expected_point"], dtype=np.uint64)
    assert oracle.pt_is_valid(expected) == 1
    r0 = oracle.fe_from_bytes(bytes.fromhex(r["elligator_input_hex"]))
    got = oracle.ris_elligator(r0)
    assert oracle.pt_is_valid(got) == 1
    assert oracle.ris_eq(got, expected) == 1
    assert oracle.ris_compress(got) == oracle.ris_compress(expected)


def test_ristretto_order_8l_point(oracle, kats):
    y = oracle.fe_from_bytes(bytes(kats["ristretto"]["order_8L_point_y_bytes"]))  # validity_check :642
    pt = oracle.pt_new_from_y_coord(y, 0)
    assert pt is not None and oracle.pt_is_valid(pt) == 1
    assert oracle.pt_eq(oracle.pt_double_and_add(pt, C(kats, "L")), oracle.pt_identity()) == 0


# ------------------------------------------------------------------------------------------------
# derived MSM oracle and batch drivers
# ------------------------------------------------------------------------------------------------
def test_msm_naive_consistency(oracle, kats):
    B = C(kats, "BASEPOINT")
    n = 24
    r = oracle.synth_scalar(4, 0, 0, n)
    s = oracle.synth_scalar(4, 1, 0, n)
    pts = oracle.pt_scalar_mul_batch(np.tile(B, (n, 1)), r, threads=4)
    got = oracle.msm_naive(pts, s)
    total = sum(pm.from_limbs(r[i]) * pm.from_limbs(s[i]) for i in range(n)) % pm.L
    want = oracle.pt_double_and_add(B, pm.to_limbs(total))
    assert oracle.pt_eq(got, want) == 1
    assert oracle.ris_compress(got) == oracle.ris_compress(want)
    got_mt = oracle.msm_naive(pts, s, threads=3)
    assert oracle.pt_eq(got_mt, got) == 1
    ones = np.tile(ONE, (n, 1))                                             # all s_i = 1 == chained Add
    acc = oracle.pt_identity()
    for i in range(n):
        acc = oracle.pt_add(acc, pts[i])
    assert oracle.pt_eq(oracle.msm_naive(pts, ones), acc) == 1


def test_batch_threads_identical(oracle):
    n = 1000
    a, b = oracle.synth_fe(5, 0, 0, n), oracle.synth_fe(5, 1, 0, n)
    assert eq(oracle.fe_mul_batch(a, b, 1), oracle.fe_mul_batch(a, b, 8))
    p, s = oracle.fe_mul_square_batch(a, b, 4)
    assert eq(p, oracle.fe_mul_batch(a, b)) and eq(s, oracle.fe_square_batch(a))
