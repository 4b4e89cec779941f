"""Independent model of the MSM plan (test infrastructure): the signed-digit recoding with and without a carry chain and
the short-window rules, written from the definitions.  tests/test_msm_plan.py and tests/test_sharding_gloo.py compare the
library's answers (dusk_zerocaf_b200.sharding -> zc_msm_plan_query) with these."""
import numpy as np

SCALAR_BITS = 250            # canonical scalars are < L < 2^250


def num_windows(window_bits):
    return (256 + window_bits - 1) // window_bits


def window_owner(w, nranks):
    """Boustrophedon over the ranks: 0..R-1, R-1..0, 0..R-1, ..."""
    q, r = divmod(w, nranks)
    return nranks - 1 - r if q & 1 else r


def signed_digits(scalar_int, window_bits):
    """d_w in [-2^(c-1), 2^(c-1)) with sum_w d_w 2^(c w) == scalar (carry-propagating recode)."""
    c = window_bits
    half, full, mask = 1 << (c - 1), 1 << c, (1 << c) - 1
    out, carry = [], 0
    for w in range(num_windows(c)):
        raw = ((scalar_int >> (c * w)) & mask) + carry
        if raw >= half:
            out.append(raw - full)
            carry = 1
        else:
            out.append(raw)
            carry = 0
    if carry:
        raise ValueError("scalar too large for the window plan (not canonical)")
    return out


def limbs_to_int(limbs):
    return sum(int(x) << (52 * i) for i, x in enumerate(np.asarray(limbs).reshape(-1)[:5]))


def offset_digits(scalar_int, window_bits):
    """msm_digits_kernel's recoding: with H = sum_w 2^(c-1) 2^(c w),  d_w = ((s + H) >> c w) mod 2^c - 2^(c-1)."""
    c = window_bits
    nwin = num_windows(c)
    H = sum(1 << (c - 1 + c * w) for w in range(nwin))
    v = scalar_int + H
    return [((v >> (c * w)) & ((1 << c) - 1)) - (1 << (c - 1)) for w in range(nwin)]


def short_window_sub_bits(window_bits, w):
    ba = max(0, SCALAR_BITS - window_bits * w)
    return (window_bits - 1) - ba if ba < window_bits - 1 else 0


def merged_spread_bits(window_bits, w):
    sub = short_window_sub_bits(window_bits, w)
    return sub - 1 if SCALAR_BITS - window_bits * w > 0 and sub > 1 else 0


def spread_digit(d, point_index, sm):
    return d * (1 << sm) + (point_index & ((1 << sm) - 1))


def fixed_base_row_shift(window_bits, w):
    return window_bits * w - merged_spread_bits(window_bits, w)
