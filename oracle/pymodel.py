"""Independent Python-bigint model of the Sonny-curve arithmetic -- TEST INFRASTRUCTURE ONLY.

A second opinion for oracle/zerocaf_oracle.c: plain `% p` arithmetic instead of the reference's
radix-2^52 Montgomery schedule.  Because every reference field op returns the canonical residue
(SURVEY.md section 0), the C oracle and this model must agree value-for-value.

Constants: /root/reference/src/backend/u64/constants.rs (:9 L, :30-36 FIELD_L, :86-92 EDWARDS_D).
"""

P = 2**252 + 27742317777372353535851937790883648493            # FIELD_L
L = 2**249 + 14490550575682688738086195780655237219            # sub-group order
D = (-126296 * pow(126297, -1, P)) % P                         # EDWARDS_D
A = P - 1                                                      # EDWARDS_A = -1
SQRT_M1 = None                                                 # filled lazily

MASK52 = (1 << 52) - 1


def to_limbs(v):
    return [(v >> (52 * i)) & MASK52 for i in range(5)]


def from_limbs(l):
    return sum(int(x) << (52 * i) for i, x in enumerate(list(l)[:5]))


def pt_from_limbs(l):
    l = [int(x) for x in l]
    return tuple(from_limbs(l[5 * k:5 * k + 5]) for k in range(4))


def pt_to_limbs(pt):
    out = []
    for c in pt:
        out += to_limbs(c)
    return out


IDENTITY = (0, 1, 1, 0)


def pt_add(p1, p2):
    """Same polynomials as reference src/edwards.rs:473-487, so (X:Y:Z:T) matches limb for limb."""
    X1, Y1, Z1, T1 = p1
    X2, Y2, Z2, T2 = p2
    a = X1 * X2 % P
    b = Y1 * Y2 % P
    c = D * T1 % P * T2 % P
    d = Z1 * Z2 % P
    e = ((X1 + Y1) * (X2 + Y2) - a - b) % P
    f = (d - c) % P
    g = (d + c) % P
    h = (b + a) % P
    return (e * f % P, g * h % P, f * g % P, e * h % P)


def pt_neg(p1):
    X, Y, Z, T = p1
    return ((-X) % P, Y, Z, (-T) % P)


def pt_double_and_add(pt, s):
    """LSB-first, reference src/edwards.rs:102-120 (limb-exact representative)."""
    n, q = pt, IDENTITY
    while s:
        if s & 1:
            q = pt_add(q, n)
        n = pt_add(n, n)
        s >>= 1
    return q


def pt_affine(pt):
    X, Y, Z, _ = pt
    zi = pow(Z, -1, P)
    return (X * zi % P, Y * zi % P)


def on_curve(pt):
    x, y = pt_affine(pt)
    return (-x * x + y * y) % P == (1 + D * x * x % P * y * y) % P


def _sqrt(a):
    """Tonelli-Shanks; returns one root or None."""
    a %= P
    if a == 0:
        return 0
    if pow(a, (P - 1) // 2, P) != 1:
        return None
    q, s = P - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    z = 2
    while pow(z, (P - 1) // 2, P) != P - 1:
        z += 1
    m, c, t, r = s, pow(z, q, P), pow(a, q, P), pow(a, (q + 1) // 2, P)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % P
            i += 1
        b = pow(c, 1 << (m - i - 1), P)
        m, c = i, b * b % P
        t, r = t * c % P, r * b % P
    return r


def is_positive(x):
    return x % P <= (P - 1) // 2


def sqrt_m1():
    global SQRT_M1
    if SQRT_M1 is None:
        r = _sqrt(P - 1)
        SQRT_M1 = r
    return SQRT_M1


def ristretto_equal(p1, p2):
    """reference src/ristretto.rs:166-176"""
    X1, Y1 = p1[0], p1[1]
    X2, Y2 = p2[0], p2[1]
    return (X1 * Y2 - Y1 * X2) % P == 0 or (X1 * X2 - Y1 * Y2) % P == 0
