"""Deterministic synthetic inputs: counter-based SplitMix64, seed 0x5A45524F43414621.

FieldElement: 32 random bytes with bytes[31] &= 0x07  -- the distribution of FieldElement::random (src/field.rs:131-140)
Scalar:       32 random bytes with bytes[31] &= 0x01  -- the distribution of Scalar::random (src/scalar.rs:100-109)
Vectorised numpy; bit-identical to the oracle's zo_synth_fe / zo_synth_scalar (checked in tests/test_synth.py).
"""
import numpy as np

SEED = 0x5A45524F43414621
_M52 = np.uint64((1 << 52) - 1)


def _splitmix(seed, stream, ctr):
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) + np.uint64(stream) * np.uint64(0xD1B54A32D192ED03)
             + (ctr + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _words(seed, stream, first, n, top_byte_mask):
    i = np.arange(first, first + n, dtype=np.uint64)
    w = [_splitmix(seed, stream, np.uint64(4) * i + np.uint64(j)) for j in range(4)]
    w[3] = w[3] & np.uint64(((top_byte_mask << 56) | ((1 << 56) - 1)))
    return w


def _limbs(w, out=None):
    n = w[0].shape[0]
    if out is None:
        out = np.empty((n, 5), dtype=np.uint64)
    u = np.uint64
    out[:, 0] = w[0] & _M52
    out[:, 1] = ((w[0] >> u(52)) | (w[1] << u(12))) & _M52
    out[:, 2] = ((w[1] >> u(40)) | (w[2] << u(24))) & _M52
    out[:, 3] = ((w[2] >> u(28)) | (w[3] << u(36))) & _M52
    out[:, 4] = w[3] >> u(16)
    return out


def synth_fe(stream, first, n, seed=SEED, out=None):
    return _limbs(_words(seed, stream, first, n, 0x07), out)


def synth_scalar(stream, first, n, seed=SEED, out=None):
    return _limbs(_words(seed, stream, first, n, 0x01), out)


# BASEPOINT (constants.rs:188-211), X|Y|Z|T radix-2^52 limbs; y = 3/5, Z = 1, T = X*Y
BASEPOINT = np.array([
    276718085098056, 1646536057461434, 2704687245600312, 2630386667454967, 13476148227069,
    1303868825475266, 3250718520537114, 2702159777242978, 2702159776422297, 10555311626649,
    1, 0, 0, 0, 0,
    3634527586288175, 2006028620404053, 3424252198034825, 2478951925947079, 4567251727358], dtype=np.uint64)
