"""Host-side mirror of the reference's value types and operator surface.

  FieldElement   /root/reference/src/backend/u64/field.rs:31-32   (alias src/field.rs:90-91)
  Scalar         /root/reference/src/backend/u64/scalar.rs:26-27  (alias src/scalar.rs:75-76)
  EdwardsPoint   /root/reference/src/edwards.rs:336-342
  RistrettoPoint /root/reference/src/ristretto.rs:157-158

Same names, same limb layout (`.limbs` == the Rust tuple field `.0`), same operators (Add/Sub/Mul/Neg, Identity,
Square, Double, Mul<Scalar>), same error behaviour where the reference panics (Scalar.from_bytes on a value > L-1).
Every operator is a 1-element call into the CUDA library -- convenient for tests that read like the reference's own;
throughput work goes through dusk_zerocaf_b200.batch.  No arithmetic is done in Python.
"""
import numpy as np

from . import batch

_MASK52 = (1 << 52) - 1
_L_INT = 2**249 + 14490550575682688738086195780655237219   # constants.rs:9


def _limbs_from_int(v):
    return np.array([(v >> (52 * i)) & _MASK52 for i in range(5)], dtype=np.uint64)


def _int_from_limbs(l):
    return sum(int(x) << (52 * i) for i, x in enumerate(l))


class _Residue:
    __slots__ = ("limbs",)
    _ops = None   # (mul, add, sub, square, neg)

    def __init__(self, limbs):
        l = np.asarray(limbs, dtype=np.uint64).reshape(5).copy()
        self.limbs = l

    # -- constructors (field.rs:513-531 / scalar.rs:330-343) --
    @classmethod
    def zero(cls): return cls([0, 0, 0, 0, 0])
    @classmethod
    def one(cls): return cls([1, 0, 0, 0, 0])
    @classmethod
    def identity(cls): return cls.one()          # Identity for FieldElement is 1 (field.rs:78-87)

    @classmethod
    def from_bytes(cls, b):
        """32 little-endian bytes -> limbs (field.rs:563-587 / scalar.rs:445-467): pure bit repacking."""
        b = bytes(b)
        if len(b) != 32:
            raise ValueError("need 32 bytes")
        return cls(_limbs_from_int(int.from_bytes(b, "little")))

    def to_bytes(self):
        return _int_from_limbs(self.limbs).to_bytes(32, "little")

    def __eq__(self, other):                      # PartialEq via to_bytes ct_eq (src/field.rs:93-106)
        return isinstance(other, type(self)) and bool(np.array_equal(self.limbs, other.limbs))

    def __hash__(self):
        return hash(self.limbs.tobytes())

    def __repr__(self):
        return f"{type(self).__name__}({[int(x) for x in self.limbs]})"

    def _b(self, k, other):
        if not isinstance(other, type(self)):
            return NotImplemented
        return type(self)(type(self)._ops[k](self.limbs, other.limbs)[0])

    def __add__(self, o): return self._b(1, o)
    def __sub__(self, o): return self._b(2, o)
    def __neg__(self): return type(self)(type(self)._ops[4](self.limbs)[0])
    def square(self): return type(self)(type(self)._ops[3](self.limbs)[0])


class FieldElement(_Residue):
    _ops = (batch.fe_mul, batch.fe_add, batch.fe_sub, batch.fe_square, batch.fe_neg)

    def __mul__(self, o): return self._b(0, o)
    def pow(self, e): return FieldElement(batch.fe_pow(self.limbs, e.limbs)[0])            # field.rs:334-354
    def half(self): return FieldElement(batch.fe_half(self.limbs)[0])                      # field.rs:317-323
    def inverse(self):                                                                     # field.rs:854-925
        if self == FieldElement.zero():
            raise ZeroDivisionError("FieldElement::inverse of zero (the reference asserts, field.rs:864)")
        return FieldElement(batch.fe_invert(self.limbs)[0])

    @staticmethod
    def sqrt_ratio_i(u, v):                                                                # field.rs:443-491
        r, sq = batch.fe_sqrt_ratio_i(u.limbs, v.limbs)
        return bool(sq[0]), FieldElement(r[0])


class Scalar(_Residue):
    _ops = (batch.scalar_mul, batch.scalar_add, batch.scalar_sub, batch.scalar_square, batch.scalar_neg)

    @classmethod
    def from_bytes(cls, b):
        v = int.from_bytes(bytes(b), "little")
        if v > _L_INT - 1:   # the reference asserts here (scalar.rs:465)
            raise ValueError("scalar out of range: value must be <= L - 1 (reference panics, scalar.rs:465)")
        return super().from_bytes(b)

    def __mul__(self, o):
        if isinstance(o, (EdwardsPoint, RistrettoPoint)):   # Mul<EdwardsPoint> for Scalar (edwards.rs:563-577)
            return o * self
        return self._b(0, o)

    def pow(self, e): return Scalar(batch.scalar_pow(self.limbs, e.limbs)[0])              # scalar.rs:293-322
    def half(self): return Scalar(batch.scalar_half(self.limbs)[0])                        # scalar.rs:285-291
    def compute_window_NAF(self, width): return batch.scalar_window_naf(self.limbs, width)[0]   # scalar.rs:396-415
    def compute_NAF(self): return self.compute_window_NAF(2)                               # scalar.rs:370-390


class EdwardsPoint:
    """Extended twisted-Edwards point; `.limbs` is X|Y|Z|T, 20 x u64 (edwards.rs:336-342)."""
    __slots__ = ("limbs",)

    def __init__(self, limbs=None, X=None, Y=None, Z=None, T=None):
        if limbs is None:
            limbs = np.concatenate([c.limbs for c in (X, Y, Z, T)])
        self.limbs = np.asarray(limbs, dtype=np.uint64).reshape(20).copy()

    X = property(lambda s: FieldElement(s.limbs[0:5]))
    Y = property(lambda s: FieldElement(s.limbs[5:10]))
    Z = property(lambda s: FieldElement(s.limbs[10:15]))
    T = property(lambda s: FieldElement(s.limbs[15:20]))

    @classmethod
    def identity(cls):                            # (0, 1, 1, 0)  edwards.rs:381-391
        l = np.zeros(20, dtype=np.uint64)
        l[5] = 1
        l[10] = 1
        return cls(l)

    def _wrap(self, limbs): return type(self)(limbs)
    def __add__(self, o):
        if not isinstance(o, type(self)): return NotImplemented
        return self._wrap(batch.point_add(self.limbs, o.limbs)[0])
    def __sub__(self, o):
        if not isinstance(o, type(self)): return NotImplemented
        return self._wrap(batch.point_sub(self.limbs, o.limbs)[0])
    def __neg__(self): return self._wrap(batch.point_neg(self.limbs)[0])
    def double(self): return self._wrap(batch.point_double(self.limbs)[0])

    def __mul__(self, s):                         # Mul<&Scalar> = double_and_add (edwards.rs:547-561)
        if not isinstance(s, Scalar): return NotImplemented
        return self._wrap(batch.point_scalar_mul(self.limbs, s.limbs, mode=0)[0])

    def __eq__(self, o):
        """Equality of group elements.  The reference compares affine coordinates (edwards.rs:360-364); the
        inversion-free equivalent X1*Z2 == X2*Z1 and Y1*Z2 == Y2*Z1 is evaluated on the device."""
        if not isinstance(o, EdwardsPoint): return False
        a = batch.fe_mul(np.stack([self.limbs[0:5], self.limbs[5:10]]), np.stack([o.limbs[10:15], o.limbs[10:15]]))
        b = batch.fe_mul(np.stack([o.limbs[0:5], o.limbs[5:10]]), np.stack([self.limbs[10:15], self.limbs[10:15]]))
        return bool(np.array_equal(a, b))

    def to_affine(self):
        """AffinePoint::from(EdwardsPoint) (edwards.rs:1085-1092): (x, y) as FieldElements."""
        xy = batch.point_to_affine(self.limbs)[0]
        return FieldElement(xy[0:5]), FieldElement(xy[5:10])

    def __hash__(self): return hash(self.limbs.tobytes())
    def __repr__(self): return f"{type(self).__name__}(X={self.X}, Y={self.Y}, Z={self.Z}, T={self.T})"


class RistrettoPoint(EdwardsPoint):
    """Newtype over EdwardsPoint (ristretto.rs:157-158): operators forward, equality is the Ristretto quotient test."""
    __slots__ = ()

    def __eq__(self, o):                          # ristretto.rs:166-176
        if not isinstance(o, RistrettoPoint): return False
        return bool(batch.ristretto_eq(self.limbs, o.limbs)[0])

    def compress(self):
        """RistrettoPoint::compress (ristretto.rs:398-425): the 32-byte CompressedRistretto."""
        return batch.ristretto_compress(self.limbs)[0].tobytes()

    @classmethod
    def decompress(cls, encoding):
        """CompressedRistretto::decompress (ristretto.rs:96-154): a RistrettoPoint, or None."""
        pts, ok = batch.ristretto_decompress(np.frombuffer(bytes(encoding), dtype=np.uint8))
        return cls(pts[0]) if ok[0] else None

    def __hash__(self): return hash(self.limbs.tobytes())
