"""Bigint model of the interleaved Montgomery squaring in csrc/zc_fe.cuh (mont_sqr_lazy): row i of the CIOS loop multiplies by
a_i the vector u = (0, .., 0, a_i, a_{i+1} << 1, (2a)_{i+2}, .., (2a)_7).  The model follows the device code's data flow word
for word (what enters each row, one reduction per row, the division by 2^32) and must give a^2 / R mod m -- for both moduli,
canonical and lazily reduced (< 4m) inputs.  CPU only: it pins the construction the generated PTX rows implement."""
import random

P = 2**252 + 27742317777372353535851937790883648493
L = 2**249 + 14490550575682688738086195780655237219
W = 1 << 32


def words(x):
    return [(x >> (32 * k)) & (W - 1) for k in range(8)]


def mont_sqr_model(a, m):
    ninv = (-pow(m, -1, W)) % W
    aw = words(a)
    d = words((2 * a) % (1 << 256))                  # words of 2a (a < 2^255)
    t = 0                                            # the sliding window as one integer
    for i in range(8):
        u = [0] * 8
        u[i] = aw[i]
        if i + 1 < 8:
            u[i + 1] = (aw[i + 1] << 1) & (W - 1)
        for k in range(i + 2, 8):
            u[k] = d[k]
        t += aw[i] * sum(u[k] << (32 * k) for k in range(8))     # the multiply row (window-relative offsets k)
        q = (t * ninv) % W
        t += q * m                                               # the reduction row
        assert t % W == 0
        t >>= 32
        assert t < (1 << 288)                                    # nine words never overflow
    return t


def test_row_vectors_sum_to_the_square():
    rng = random.Random(1)
    for _ in range(200):
        a = rng.getrandbits(255)
        aw = words(a)
        d = words(2 * a)
        total = 0
        for i in range(8):
            u = (aw[i] << (32 * i)) + (((aw[i + 1] << 1) & (W - 1)) << (32 * (i + 1)) if i + 1 < 8 else 0) \
                + sum(d[k] << (32 * k) for k in range(i + 2, 8))
            total += (aw[i] << (32 * i)) * u
        assert total == a * a


def test_model_is_a_montgomery_squaring():
    rng = random.Random(2)
    for m in (P, L):
        rinv = pow(1 << 256, -1, m)
        edge = [0, 1, m - 1, m, 2 * m - 1, 4 * m - 1, (1 << 252) - 1, W - 1, (W - 1) << 224]
        for a in edge + [rng.randrange(4 * m) for _ in range(500)]:
            r = mont_sqr_model(a, m)
            assert r % m == (a * a * rinv) % m
            assert r < a * a // (1 << 256) + m + 1              # (a^2 + Q m) / R < a^2 / R + m
            if a < 2 * m:
                assert r < 2 * m                                # one conditional subtraction makes it canonical
