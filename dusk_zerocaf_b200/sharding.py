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
