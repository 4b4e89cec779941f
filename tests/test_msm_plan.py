"""Host-side invariants of the MSM plan that csrc/zc_msm.cu implements (asked from the library through dusk_zerocaf_b200/sharding.py -> zc_msm_plan_query, compared with the independent model tests/msm_plan_model.py):
the carry-free digit recoding, the short-window sub-bucket rule, and the spread rule + constant correction of the
fixed-base (merged bucket set) path.  Pure integer checks, no GPU."""
import numpy as np

from dusk_zerocaf_b200 import sharding
import msm_plan_model as model

L = 2**249 + 14490550575682688738086195780655237219


def _scalars(rng, k=200):
    edge = [0, 1, 2, L - 1, L - 2, 2**249 - 1, 2**249, 2**248, (1 << 249) + (1 << 100), 0x8888888888888888 << 180]
    for c in range(8, 17):
        edge += [2**(c - 1), 2**(c - 1) - 1, 2**c - 1, (2**c - 1) << (c * 3), ((1 << 249) - 1) ^ (1 << (c * 2))]
    rnd = [int.from_bytes(rng.bytes(32), "little") % L for _ in range(k)]
    return [s % L for s in edge] + rnd


def test_offset_recoding_equals_carry_recoding():
    rng = np.random.default_rng(11)
    for c in range(8, 17):
        for s in _scalars(rng):
            d = model.offset_digits(s, c)
            assert d == model.signed_digits(s, c), (c, s)
            assert sum(x << (c * w) for w, x in enumerate(d)) == s


def test_short_windows_only_see_small_nonnegative_digits():
    """The premise of both short-window rules: for a canonical scalar a window starting at bit c w > 250 - c holds a digit in
    [0, 2^(250 - c w - 1) + 1]; a window starting exactly at bit 250 can still receive a carry (digit 0 or 1: c = 10, scalars
    >= 2^249), windows above hold 0."""
    rng = np.random.default_rng(12)
    for c in range(8, 17):
        for s in _scalars(rng, 400):
            d = model.offset_digits(s, c)
            for w, x in enumerate(d):
                ba = 250 - c * w
                if ba == 0:
                    assert x in (0, 1), (c, w, s)
                elif ba < 0:
                    assert x == 0, (c, w, s)
                elif ba < c - 1:
                    assert 0 <= x <= (1 << (ba - 1)) + 1, (c, w, s, x)


def test_sub_buckets_fit_the_bucket_range():
    for c in range(8, 17):
        nb = 1 << (c - 1)
        for w in range(sharding.num_windows(c)):
            sub = sharding.short_window_sub_bits(c, w)
            ba = max(0, 250 - c * w)
            if sub:
                # slot = ((d - 1) << sub) | low bits of the point index, d <= 2^ba
                assert (((1 << ba) - 1) << sub | ((1 << sub) - 1)) < nb, (c, w)


def test_fixed_base_spread_weights_fit_and_sum_to_the_scalar():
    """Entry weight d' = d 2^SM + (i mod 2^SM) stays inside [0, 2^(c-1)] for every canonical scalar, and
    sum_w weight_w 2^shift_w  -  sum over spread windows of (i mod 2^SM) 2^shift_w  ==  s  (the constant the chain subtracts)."""
    rng = np.random.default_rng(13)
    for c in range(8, 17):
        nwin = sharding.num_windows(c)
        sms = [sharding.merged_spread_bits(c, w) for w in range(nwin)]
        assert sum(1 for x in sms if x) <= 1                     # at most one spread window per scalar width
        for s in _scalars(rng, 300):
            d = model.offset_digits(s, c)
            for i in (0, 1, 5, 2**20 - 1, 12345):
                total, corr = 0, 0
                for w in range(nwin):
                    sh = sharding.fixed_base_row_shift(c, w)
                    if sms[w]:
                        wgt = model.spread_digit(d[w], i, sms[w])
                        assert 0 <= wgt <= (1 << (c - 1)), (c, w, s, i, wgt)
                        corr += (i & ((1 << sms[w]) - 1)) << sh
                        assert sh + sms[w] == c * w and ((i & ((1 << sms[w]) - 1)) << sh) < L   # the correction scalars are canonical
                    else:
                        wgt = d[w]
                        assert -(1 << (c - 1)) <= wgt < (1 << (c - 1))
                    total += wgt << sh if wgt >= 0 else -((-wgt) << sh)
                assert total - corr == s, (c, s, i)


def test_fixed_base_table_size():
    assert sharding.fixed_base_table_rows(16, 0, 1, 1 << 20) * 128 == 2 << 30          # 2 GiB on one GPU
    assert all(sharding.fixed_base_table_rows(16, r, 8, 1 << 20) * 128 == 256 << 20 for r in range(8))
    assert sharding.fixed_base_table_rows(16, 20, 32, 1000) == 0                        # a rank beyond the window count
    assert sharding.fixed_base_row_shift(16, 15) == 236 and sharding.merged_spread_bits(16, 15) == 4
    assert sharding.short_window_sub_bits(16, 15) == 5 and sharding.short_window_sub_bits(16, 14) == 0


def test_library_plan_rules_equal_the_model():
    """zc_msm_plan_query (the C side) against the model, every window size and window, several rank counts."""
    for c in range(8, 17):
        for w in range(model.num_windows(c)):
            assert sharding.short_window_sub_bits(c, w) == model.short_window_sub_bits(c, w), (c, w)
            assert sharding.merged_spread_bits(c, w) == model.merged_spread_bits(c, w), (c, w)
            assert sharding.fixed_base_row_shift(c, w) == model.fixed_base_row_shift(c, w), (c, w)
            for R in (1, 2, 3, 4, 8, 16, 32):
                assert sharding.plan_query(c, w, R)[0] == model.window_owner(w, R), (c, w, R)
    # the pairing the boustrophedon order buys at 8 ranks, 16 windows: rank r owns windows r and 15 - r
    assert [sharding.windows_of_rank(16, r, 8) for r in range(8)] == [[r, 15 - r] for r in range(8)]
    assert sharding.windows_of_rank(16, 0, 4) == [0, 7, 8, 15]
