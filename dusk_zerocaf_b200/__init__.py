"""dusk_zerocaf_b200 -- B200-native (sm_100a) batched backend for the arithmetic hot path of dusk-zerocaf.

Python is only the host-side mirror of the reference's type/trait surface (FieldElement, Scalar, EdwardsPoint,
RistrettoPoint with Add/Sub/Mul/Neg/Identity/Double/Square) plus the batch API; all arithmetic runs in hand-written
CUDA kernels behind the C ABI of libzerocaf_b200.so (include/zerocaf_b200.h).  There is no CPU fallback.
"""
from ._lib import ZerocafError, SO_PATH, header_symbols, build, lib   # noqa: F401
from .context import Context, Generators, default_context, GEN_PREPARED, GEN_FIXED_BASE   # noqa: F401
from .types import FieldElement, Scalar, EdwardsPoint, RistrettoPoint   # noqa: F401
from . import batch, synth   # noqa: F401

SCALAR_MUL_STRICT = 0
SCALAR_MUL_FAST = 1
