"""Bigint model of batch_invert in csrc/zc_encode.cu (Montgomery's trick on the loaded normal-form words).  The device code
never converts its inputs: a normal-form value a, read as a Montgomery-form word, stands for a / R; the chain therefore
produces R^2 / a and SCALE products by 1 on the running inverse turn that into 1 / a (SCALE 2: fe_invert) or R / a (SCALE 1:
Div and extended -> affine multiply a normal-form value by it).  Zeros are skipped in the product and come back as zero."""
import random

P = 2**252 + 27742317777372353535851937790883648493
R = 1 << 256
RINV = pow(R, -1, P)


def mm(x, y):                      # what mont_mul computes on words
    return (x * y * RINV) % P


def mont_pow(x, e):                # fe_pow_const: left-to-right square-and-multiply with Montgomery products
    r = x
    for bit in range(e.bit_length() - 2, -1, -1):
        r = mm(r, r)
        if (e >> bit) & 1:
            r = mm(r, x)
    return r


def batch_invert_model(vals, scale):
    one = R % P
    pref, acc = [], one
    for x in vals:
        pref.append(acc)
        if x:
            acc = mm(acc, x)
    inv = mont_pow(acc, P - 2)
    for _ in range(scale):
        inv = mm(inv, 1)
    out = [0] * len(vals)
    for j in range(len(vals) - 1, -1, -1):
        r = mm(inv, pref[j])
        if vals[j]:
            inv = mm(inv, vals[j])
            out[j] = r
    return out


def test_scale_2_is_the_plain_inverse_and_scale_1_its_montgomery_form():
    rng = random.Random(3)
    for _ in range(50):
        vals = [rng.randrange(P) for _ in range(8)]
        for z in rng.sample(range(8), rng.randrange(0, 4)):
            vals[z] = 0
        inv2 = batch_invert_model(vals, 2)
        inv1 = batch_invert_model(vals, 1)
        for a, i2, i1 in zip(vals, inv2, inv1):
            if a == 0:
                assert i2 == 0 and i1 == 0                    # inverse(0) = 0 (the reference panics, field.rs:864)
            else:
                assert (a * i2) % P == 1
                assert i1 == (i2 * R) % P
                b = rng.randrange(P)
                assert mm(b, i1) == (b * pow(a, -1, P)) % P    # Div: normal-form b times R / a -> b / a
