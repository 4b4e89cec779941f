"""Batch API over host numpy arrays (uint64; FieldElement/Scalar = (n,5), points = (n,20)).

Each function is one call through the C ABI with HOST buffers (copy in, kernel, copy out).  Names follow the
reference's operators: field.rs / scalar.rs / edwards.rs / ristretto.rs Add, Sub, Mul, Neg, Square, Double.
"""
import numpy as np

from .context import default_context


def _arr(x, cols):
    x = np.ascontiguousarray(x, dtype=np.uint64)
    if x.ndim == 1:
        x = x.reshape(1, -1)
    if x.shape[-1] != cols:
        raise ValueError(f"expected (..., {cols}) uint64 limbs, got {x.shape}")
    return x


def _bin(name, a, b, cols, ctx, out_cols=None):
    ctx = ctx or default_context()
    a, b = _arr(a, cols), _arr(b, cols)
    if a.shape != b.shape:
        raise ValueError("operand shapes differ")
    out = np.empty((a.shape[0], out_cols or cols), dtype=np.uint64)
    ctx.call(name, a, b, out, a.shape[0])
    return out


def _un(name, a, cols, ctx):
    ctx = ctx or default_context()
    a = _arr(a, cols)
    out = np.empty_like(a)
    ctx.call(name, a, out, a.shape[0])
    return out


def fe_mul(a, b, ctx=None): return _bin("zc_fe_mul_batch", a, b, 5, ctx)
def fe_add(a, b, ctx=None): return _bin("zc_fe_add_batch", a, b, 5, ctx)
def fe_sub(a, b, ctx=None): return _bin("zc_fe_sub_batch", a, b, 5, ctx)
def fe_square(a, ctx=None): return _un("zc_fe_square_batch", a, 5, ctx)
def fe_neg(a, ctx=None): return _un("zc_fe_neg_batch", a, 5, ctx)
def scalar_mul(a, b, ctx=None): return _bin("zc_scalar_mul_batch", a, b, 5, ctx)
def scalar_add(a, b, ctx=None): return _bin("zc_scalar_add_batch", a, b, 5, ctx)
def scalar_sub(a, b, ctx=None): return _bin("zc_scalar_sub_batch", a, b, 5, ctx)
def scalar_square(a, ctx=None): return _un("zc_scalar_square_batch", a, 5, ctx)
def scalar_neg(a, ctx=None): return _un("zc_scalar_neg_batch", a, 5, ctx)
def point_add(p, q, ctx=None): return _bin("zc_point_add_batch", p, q, 20, ctx)
def point_sub(p, q, ctx=None): return _bin("zc_point_sub_batch", p, q, 20, ctx)
def point_double(p, ctx=None): return _un("zc_point_double_batch", p, 20, ctx)
def point_neg(p, ctx=None): return _un("zc_point_neg_batch", p, 20, ctx)


def fe_mul_square(a, b, ctx=None):
    ctx = ctx or default_context()
    a, b = _arr(a, 5), _arr(b, 5)
    prod, sq = np.empty_like(a), np.empty_like(a)
    ctx.call("zc_fe_mul_square_batch", a, b, prod, sq, a.shape[0])
    return prod, sq


def point_scalar_mul(points, scalars, mode=0, ctx=None):
    ctx = ctx or default_context()
    p, s = _arr(points, 20), _arr(scalars, 5)
    if p.shape[0] != s.shape[0]:
        raise ValueError("points / scalars length differ")
    out = np.empty_like(p)
    ctx.call("zc_point_scalar_mul_batch", p, s, out, p.shape[0], int(mode))
    return out


def basepoint_mul(scalars, ctx=None):
    """[s_i] B for the basepoint (constants.rs:188-211) through the fixed-base table; compare canonically."""
    ctx = ctx or default_context()
    s = _arr(scalars, 5)
    out = np.empty((s.shape[0], 20), dtype=np.uint64)
    ctx.call("zc_basepoint_mul_batch", s, out, s.shape[0])
    return out


def ristretto_eq(p, q, ctx=None):
    ctx = ctx or default_context()
    p, q = _arr(p, 20), _arr(q, 20)
    out = np.empty(p.shape[0], dtype=np.uint8)
    ctx.call("zc_ristretto_eq_batch", p, q, out, p.shape[0])
    return out


def fe_invert(a, ctx=None):
    """FieldElement::inverse (field.rs:854-925); inverse(0) = 0 here, the reference panics."""
    return _un("zc_fe_invert_batch", a, 5, ctx)


def point_to_affine(p, ctx=None):
    """AffinePoint::from(EdwardsPoint) (edwards.rs:1085-1092): (n, 10) limbs = x | y."""
    ctx = ctx or default_context()
    p = _arr(p, 20)
    out = np.empty((p.shape[0], 10), dtype=np.uint64)
    ctx.call("zc_point_to_affine_batch", p, out, p.shape[0])
    return out


def ristretto_compress(p, ctx=None):
    """RistrettoPoint::compress (ristretto.rs:398-425): (n, 32) uint8."""
    ctx = ctx or default_context()
    p = _arr(p, 20)
    out = np.empty((p.shape[0], 32), dtype=np.uint8)
    ctx.call("zc_ristretto_compress_batch", p, out, p.shape[0])
    return out


def ristretto_decompress(encodings, ctx=None):
    """CompressedRistretto::decompress (ristretto.rs:96-154): ((n, 20) points, (n,) ok flags); ok == 0 <=> None."""
    ctx = ctx or default_context()
    e = np.ascontiguousarray(encodings, dtype=np.uint8).reshape(-1, 32)
    out = np.empty((e.shape[0], 20), dtype=np.uint64)
    ok = np.empty(e.shape[0], dtype=np.uint8)
    ctx.call("zc_ristretto_decompress_batch", e, out, ok, e.shape[0])
    return out, ok


def ristretto_elligator(r0, ctx=None):
    """RistrettoPoint::elligator_ristretto_flavor (ristretto.rs:430-471), limb-exact: (n, 20)."""
    ctx = ctx or default_context()
    r0 = _arr(r0, 5)
    out = np.empty((r0.shape[0], 20), dtype=np.uint64)
    ctx.call("zc_ristretto_elligator_batch", r0, out, r0.shape[0])
    return out


def ristretto_from_uniform_bytes(data, ctx=None):
    """RistrettoPoint::from_uniform_bytes (ristretto.rs:493-507): (n, 64) uint8 -> (n, 20), limb-exact."""
    ctx = ctx or default_context()
    d = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1, 64)
    out = np.empty((d.shape[0], 20), dtype=np.uint64)
    ctx.call("zc_ristretto_from_uniform_bytes_batch", d, out, d.shape[0])
    return out


def point_is_valid(p, ctx=None):
    """ValidityCheck for EdwardsPoint (edwards.rs:393-400): the curve equation in projective coordinates."""
    ctx = ctx or default_context()
    p = _arr(p, 20)
    ok = np.empty(p.shape[0], dtype=np.uint8)
    ctx.call("zc_point_is_valid_batch", p, ok, p.shape[0])
    return ok


def fe_pow(a, e, ctx=None):
    """Pow for FieldElement (field.rs:334-354): a^e mod p, 0^0 = 1."""
    return _bin("zc_fe_pow_batch", a, e, 5, ctx)


def scalar_pow(a, e, ctx=None):
    """Pow for Scalar (scalar.rs:293-322): a^e mod L."""
    return _bin("zc_scalar_pow_batch", a, e, 5, ctx)


def fe_half(a, ctx=None):
    """Half for FieldElement (field.rs:317-323): a * 2^-1 mod p."""
    return _un("zc_fe_half_batch", a, 5, ctx)


def scalar_half(a, ctx=None):
    """Half for Scalar (scalar.rs:285-291): a * 2^-1 mod L."""
    return _un("zc_scalar_half_batch", a, 5, ctx)


def _to_bytes(name, a, ctx):
    ctx = ctx or default_context()
    a = _arr(a, 5)
    out = np.empty((a.shape[0], 32), dtype=np.uint8)
    ctx.call(name, a, out, a.shape[0])
    return out


def fe_to_bytes(a, ctx=None):
    """FieldElement::to_bytes (field.rs:591-631): (n, 32) uint8."""
    return _to_bytes("zc_fe_to_bytes_batch", a, ctx)


def scalar_to_bytes(a, ctx=None):
    """Scalar::to_bytes (scalar.rs:477-516): (n, 32) uint8."""
    return _to_bytes("zc_scalar_to_bytes_batch", a, ctx)


def fe_from_bytes(data, ctx=None):
    """FieldElement::from_bytes (field.rs:563-587): all 256 bits kept, no reduction."""
    ctx = ctx or default_context()
    d = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1, 32)
    out = np.empty((d.shape[0], 5), dtype=np.uint64)
    ctx.call("zc_fe_from_bytes_batch", d, out, d.shape[0])
    return out


def scalar_from_bytes(data, ctx=None):
    """Scalar::from_bytes (scalar.rs:445-467): ((n, 5) limbs, (n,) ok); ok == 0 where the reference asserts (value >= L)."""
    ctx = ctx or default_context()
    d = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1, 32)
    out = np.empty((d.shape[0], 5), dtype=np.uint64)
    ok = np.empty(d.shape[0], dtype=np.uint8)
    ctx.call("zc_scalar_from_bytes_batch", d, out, ok, d.shape[0])
    return out, ok


def scalar_window_naf(a, width, ctx=None):
    """Scalar::compute_window_NAF (scalar.rs:396-415; width 2 = compute_NAF :370-390): (n, 256) int8, LSB first."""
    ctx = ctx or default_context()
    a = _arr(a, 5)
    out = np.empty((a.shape[0], 256), dtype=np.int8)
    ctx.check(ctx._L.zc_scalar_window_naf_batch(ctx._h, a.ctypes.data, int(width), out.ctypes.data, a.shape[0]))
    return out


def fe_sqrt_ratio_i(u, v, ctx=None):
    """FieldElement::sqrt_ratio_i (field.rs:443-491): ((n, 5) root, (n,) was_square)."""
    ctx = ctx or default_context()
    u, v = _arr(u, 5), _arr(v, 5)
    out = np.empty((u.shape[0], 5), dtype=np.uint64)
    sq = np.empty(u.shape[0], dtype=np.uint8)
    ctx.call("zc_fe_sqrt_ratio_i_batch", u, v, out, sq, u.shape[0])
    return out, sq


def fe_div(a, b, ctx=None):
    """Div (field.rs:277-299): a * b^-1; a / 0 = 0 where the reference asserts."""
    ctx = ctx or default_context()
    a, b = _arr(a, 5), _arr(b, 5)
    out = np.empty_like(a)
    ctx.call("zc_fe_div_batch", a, b, out, a.shape[0])
    return out


def scalar_into_bits(a, ctx=None):
    """Scalar::into_bits (scalar.rs:352-366): (n, 256) uint8, least significant bit first."""
    ctx = ctx or default_context()
    a = _arr(a, 5)
    out = np.empty((a.shape[0], 256), dtype=np.uint8)
    ctx.call("zc_scalar_into_bits_batch", a, out, a.shape[0])
    return out


def fe_mul_square_packed(a_bytes, b_bytes, ctx=None):
    """Config 2 on the 32-byte wire format (to_bytes / from_bytes, field.rs:563-631): ((n, 32) prod, (n, 32) sq) uint8."""
    ctx = ctx or default_context()
    a = np.ascontiguousarray(a_bytes, dtype=np.uint8).reshape(-1, 32)
    b = np.ascontiguousarray(b_bytes, dtype=np.uint8).reshape(-1, 32)
    prod, sq = np.empty_like(a), np.empty_like(a)
    ctx.call("zc_fe_mul_square_batch_packed", a, b, prod, sq, a.shape[0])
    return prod, sq


def check_canonical(kind, a, ctx=None):
    """zc_{fe,scalar,point}_check_canonical_batch: None when every element is canonical, else the first offending index."""
    import ctypes
    ctx = ctx or default_context()
    a = _arr(a, 20 if kind == "point" else 5)
    bad = ctypes.c_uint64(0)
    st = getattr(ctx._L, f"zc_{kind}_check_canonical_batch")(ctx._h, a.ctypes.data, a.shape[0], ctypes.byref(bad))
    if st == 4:
        return int(bad.value)
    ctx.check(st)
    return None


def msm(points, scalars, window_bits=16, ctx=None):
    """sum_i [s_i] P_i as an EdwardsPoint (20 limbs); a group element, compare canonically."""
    ctx = ctx or default_context()
    out = np.empty(20, dtype=np.uint64)
    if points is None or len(points) == 0:
        ctx.check(ctx._L.zc_msm(ctx._h, None, None, 0, int(window_bits), out.ctypes.data))
        return out
    p, s = _arr(points, 20), _arr(scalars, 5)
    if p.shape[0] != s.shape[0]:
        raise ValueError("points / scalars length differ")
    ctx.check(ctx._L.zc_msm(ctx._h, p.ctypes.data, s.ctypes.data, p.shape[0], int(window_bits), out.ctypes.data))
    return out
