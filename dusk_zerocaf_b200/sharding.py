"""The MSM sharding plan, asked from the library (zc_msm_plan_query in csrc/zc_msm.cu -- host code, no device needed):
window ownership per rank and the short-window rules.  Nothing is restated here; tests/msm_plan_model.py holds the
independent model the CPU tests compare these answers (and the digit recoding) with."""
import ctypes

from ._lib import lib


def num_windows(window_bits):
    """ceil(256 / c): canonical scalars are < L < 2^250, so the top window never overflows for c in 8..16."""
    if not 8 <= window_bits <= 16:
        raise ValueError("window_bits must be in 8..16")
    return (256 + window_bits - 1) // window_bits


def plan_query(window_bits, window, nranks):
    """(owner rank, sub-bucket bits, merged spread bits, fixed-base row shift) of a window."""
    out = (ctypes.c_int32 * 4)()
    st = lib().zc_msm_plan_query(int(window_bits), int(window), int(nranks), out)
    if st != 0:
        raise ValueError(f"zc_msm_plan_query status {st}")
    return tuple(int(x) for x in out)


def windows_of_rank(window_bits, rank, nranks):
    if nranks < 1 or not 0 <= rank < nranks:
        raise ValueError("bad rank / nranks")
    return [w for w in range(num_windows(window_bits)) if plan_query(window_bits, w, nranks)[0] == rank]


def tasks_of_rank(window_bits, rank, nranks, n):
    """The (window, p0, p1) tasks of a rank, ascending by window: the windows it owns over all n points."""
    return [(w, 0, n) for w in windows_of_rank(window_bits, rank, nranks)]


def short_window_sub_bits(window_bits, w):
    return plan_query(window_bits, w, 1)[1]


def merged_spread_bits(window_bits, w):
    return plan_query(window_bits, w, 1)[2]


def fixed_base_row_shift(window_bits, w):
    return plan_query(window_bits, w, 1)[3]


def fixed_base_table_rows(window_bits, rank, nranks, n):
    """Rows (128 bytes each) of the fixed-base tables a ZC_GEN_FIXED_BASE handle builds on a rank."""
    return len(windows_of_rank(window_bits, rank, nranks)) * n
