"""The product's numpy input generator is bit-identical to the oracle's (same synthetic arrays on CPU and GPU legs)."""
import numpy as np

from conftest import SEED, kat_arr
from dusk_zerocaf_b200 import synth


def test_synth_matches_oracle(oracle):
    for stream, first, n in ((1, 0, 1000), (9, 12345, 257), (2, 2**24 - 5, 5)):
        assert np.array_equal(synth.synth_fe(stream, first, n), oracle.synth_fe(SEED, stream, first, n))
        assert np.array_equal(synth.synth_scalar(stream, first, n), oracle.synth_scalar(SEED, stream, first, n))


def test_synth_ranges():
    fe = synth.synth_fe(3, 0, 4096)
    sc = synth.synth_scalar(3, 0, 4096)
    assert int(fe[:, :4].max()) < 2**52 and int(fe[:, 4].max()) < 2**43    # < 2^251
    assert int(sc[:, 4].max()) < 2**41                                     # < 2^249


def test_basepoint_constant(kats):
    assert np.array_equal(synth.BASEPOINT, kat_arr(kats, "constants", "BASEPOINT"))
