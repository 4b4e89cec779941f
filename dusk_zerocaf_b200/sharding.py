"""Host-side statement of the MSM sharding plan that csrc/zc_msm.cu implements (window ownership and signed digits).

Used to size work per rank and by the CPU (gloo) tests of the multi-GPU decomposition; no point arithmetic here.
"""
import numpy as np


def num_windows(window_bits):
    """ceil(256 / c): canonical scalars are < L < 2^250, so the top window never overflows for c in 8..16."""
    if not 8 <= window_bits <= 16:
        raise ValueError("window_bits must be in 8..16")
    return (256 + window_bits - 1) // window_bits


def windows_of_rank(window_bits, rank, nranks):
    """Bucket-window sharding: window w belongs to rank w mod nranks."""
    if nranks < 1 or not 0 <= rank < nranks:
        raise ValueError("bad rank / nranks")
    return [w for w in range(num_windows(window_bits)) if w % nranks == rank]


def tasks_of_rank(window_bits, rank, nranks, n):
    """The (window, p0, p1) tasks of a rank, ascending by window -- the plan csrc/zc_msm.cu builds: the windows
    w = rank (mod nranks) over all n points.  Every (window, point) pair is owned by exactly one rank."""
    return [(w, 0, n) for w in windows_of_rank(window_bits, rank, nranks)]


def signed_digits(scalar_int, window_bits):
    """d_w in [-2^(c-1), 2^(c-1)) with sum_w d_w 2^(c w) == scalar (same recoding as msm_digits_kernel)."""
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


# ---- the same digits without a carry chain, and the short-window rules (csrc/zc_msm.cu) -------------------------------
SCALAR_BITS = 250            # canonical scalars are < L < 2^250


def offset_digits(scalar_int, window_bits):
    """msm_digits_kernel's recoding: with H = sum_w 2^(c-1) 2^(c w),  d_w = ((s + H) >> c w) mod 2^c - 2^(c-1).
    Identical to signed_digits() for every canonical scalar."""
    c = window_bits
    nwin = num_windows(c)
    H = sum(1 << (c - 1 + c * w) for w in range(nwin))
    v = scalar_int + H
    return [((v >> (c * w)) & ((1 << c) - 1)) - (1 << (c - 1)) for w in range(nwin)]


def short_window_sub_bits(window_bits, w):
    """A window that starts at bit c w >= 250 - (c-1) only sees digits in [0, 2^(250 - c w)]: the plain path spreads each of
    its digits over 2^SUB sub-buckets chosen by the point index."""
    ba = max(0, SCALAR_BITS - window_bits * w)
    return (window_bits - 1) - ba if ba < window_bits - 1 else 0


def merged_spread_bits(window_bits, w):
    """Fixed-base (merged bucket set) path: rows of a short window are scaled by 2^(c w - SM) and its digit becomes
    d 2^SM + (i mod 2^SM)."""
    sub = short_window_sub_bits(window_bits, w)
    return sub - 1 if SCALAR_BITS - window_bits * w > 0 and sub > 1 else 0


def spread_digit(d, point_index, sm):
    """The entry weight of point i in a spread short window (>= 0 for canonical scalars)."""
    return d * (1 << sm) + (point_index & ((1 << sm) - 1))


def fixed_base_table_rows(window_bits, rank, nranks, n):
    """Rows (128 bytes each) of the tables zc_msm_prepare_fixed_base_dev builds on a rank."""
    return len(windows_of_rank(window_bits, rank, nranks)) * n


def fixed_base_row_shift(window_bits, w):
    """A table row of window w is 2^shift * P_i."""
    return window_bits * w - merged_spread_bits(window_bits, w)
