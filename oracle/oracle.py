"""ctypes binding of oracle/libzerocaf_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  The product package (dusk_zerocaf_b200) never does.

All arrays are numpy uint64, C-contiguous: FieldElement / Scalar = (..., 5), EdwardsPoint = (..., 20).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libzerocaf_oracle.so")

_u64p = ctypes.POINTER(ctypes.c_uint64)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_i8p = ctypes.POINTER(ctypes.c_int8)


def build(force=False):
    """Compile the oracle with oracle/Makefile (gcc). Building the checker is not using it."""
    src = os.path.join(_HERE, "zerocaf_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _a(x, cols=None):
    x = np.ascontiguousarray(x, dtype=np.uint64)
    if cols is not None:
        assert x.shape[-1] == cols, (x.shape, cols)
    return x


def _p(x):
    return x.ctypes.data_as(_u64p)


def _bytes_in(b):
    b = np.ascontiguousarray(np.frombuffer(bytes(b), dtype=np.uint8) if not isinstance(b, np.ndarray) else b, dtype=np.uint8)
    return b


# ---- generic callers ---------------------------------------------------------------------------
def _un(name, a, cols_in=5, cols_out=5):
    a = _a(a, cols_in)
    out = np.zeros(cols_out, dtype=np.uint64)
    getattr(lib(), name)(_p(a), _p(out))
    return out


def _bin(name, a, b, ca=5, cb=5, co=5):
    a, b = _a(a, ca), _a(b, cb)
    out = np.zeros(co, dtype=np.uint64)
    getattr(lib(), name)(_p(a), _p(b), _p(out))
    return out


# ---- FieldElement --------------------------------------------------------------------------------
def fe_add(a, b): return _bin("zo_fe_add", a, b)
def fe_sub(a, b): return _bin("zo_fe_sub", a, b)
def fe_mul(a, b): return _bin("zo_fe_mul", a, b)
def fe_neg(a): return _un("zo_fe_neg", a)
def fe_square(a): return _un("zo_fe_square", a)
def fe_montgomery_mul(a, b): return _bin("zo_fe_montgomery_mul", a, b)
def fe_to_montgomery(a): return _un("zo_fe_to_montgomery", a)
def fe_from_montgomery(a): return _un("zo_fe_from_montgomery", a)
def fe_half_without_mod(a): return _un("zo_fe_half_without_mod", a)
def fe_half(a): return _un("zo_fe_half", a)
def fe_pow(a, e): return _bin("zo_fe_pow", a, e)


def fe_two_pow_k(k):
    out = np.zeros(5, dtype=np.uint64)
    lib().zo_fe_two_pow_k(ctypes.c_uint64(k), _p(out))
    return out


def fe_from_bytes(b):
    b = _bytes_in(b)
    assert b.size == 32
    out = np.zeros(5, dtype=np.uint64)
    lib().zo_fe_from_bytes(b.ctypes.data_as(_u8p), _p(out))
    return out


def fe_to_bytes(a):
    a = _a(a, 5)
    out = np.zeros(32, dtype=np.uint8)
    lib().zo_fe_to_bytes(_p(a), out.ctypes.data_as(_u8p))
    return out.tobytes()


def fe_inverse(a):
    a = _a(a, 5)
    out = np.zeros(5, dtype=np.uint64)
    if lib().zo_fe_inverse(_p(a), _p(out)):
        raise ZeroDivisionError("inverse of zero (reference panics, field.rs:864)")
    return out


def fe_div(a, b):
    a, b = _a(a, 5), _a(b, 5)
    out = np.zeros(5, dtype=np.uint64)
    if lib().zo_fe_div(_p(a), _p(b), _p(out)):
        raise ZeroDivisionError("Cannot divide by zero. (field.rs:285)")
    return out


def fe_legendre_symbol(a): return int(lib().zo_fe_legendre_symbol(_p(_a(a, 5))))
def fe_is_positive(a): return int(lib().zo_fe_is_positive(_p(_a(a, 5))))
def fe_cmp(a, b): return int(lib().zo_fe_cmp(_p(_a(a, 5)), _p(_a(b, 5))))


def fe_mod_sqrt(a, sign):
    out = np.zeros(5, dtype=np.uint64)
    ok = lib().zo_fe_mod_sqrt(_p(_a(a, 5)), int(sign), _p(out))
    return out if ok else None


def fe_sqrt_ratio_i(u, v):
    out = np.zeros(5, dtype=np.uint64)
    c = lib().zo_fe_sqrt_ratio_i(_p(_a(u, 5)), _p(_a(v, 5)), _p(out))
    return int(c), out


def fe_inv_sqrt(a):
    out = np.zeros(5, dtype=np.uint64)
    c = lib().zo_fe_inv_sqrt(_p(_a(a, 5)), _p(out))
    return int(c), out


# ---- Scalar ----------------------------------------------------------------------------------------
def sc_add(a, b): return _bin("zo_sc_add", a, b)
def sc_sub(a, b): return _bin("zo_sc_sub", a, b)
def sc_mul(a, b): return _bin("zo_sc_mul", a, b)
def sc_neg(a): return _un("zo_sc_neg", a)
def sc_square(a): return _un("zo_sc_square", a)
def sc_montgomery_mul(a, b): return _bin("zo_sc_montgomery_mul", a, b)
def sc_to_montgomery(a): return _un("zo_sc_to_montgomery", a)
def sc_from_montgomery(a): return _un("zo_sc_from_montgomery", a)
def sc_half_without_mod(a): return _un("zo_sc_half_without_mod", a)
def sc_half(a): return _un("zo_sc_half", a)
def sc_pow(a, e): return _bin("zo_sc_pow", a, e)


def sc_shr(a, n):
    out = np.zeros(5, dtype=np.uint64)
    lib().zo_sc_shr(_p(_a(a, 5)), ctypes.c_uint(n), _p(out))
    return out


def sc_two_pow_k(k):
    out = np.zeros(5, dtype=np.uint64)
    lib().zo_sc_two_pow_k(ctypes.c_uint64(k), _p(out))
    return out


def sc_from_i8(x):
    out = np.zeros(5, dtype=np.uint64)
    lib().zo_sc_from_i8(ctypes.c_int8(x), _p(out))
    return out


def sc_from_bytes(b):
    b = _bytes_in(b)
    assert b.size == 32
    out = np.zeros(5, dtype=np.uint64)
    if lib().zo_sc_from_bytes(b.ctypes.data_as(_u8p), _p(out)):
        raise ValueError("scalar > L-1 (reference asserts, scalar.rs:465)")
    return out


def sc_to_bytes(a):
    out = np.zeros(32, dtype=np.uint8)
    lib().zo_sc_to_bytes(_p(_a(a, 5)), out.ctypes.data_as(_u8p))
    return out.tobytes()


def sc_into_bits(a):
    out = np.zeros(256, dtype=np.uint8)
    lib().zo_sc_into_bits(_p(_a(a, 5)), out.ctypes.data_as(_u8p))
    return out


def sc_compute_naf(a):
    out = np.zeros(256, dtype=np.int8)
    lib().zo_sc_compute_naf(_p(_a(a, 5)), out.ctypes.data_as(_i8p))
    return out


def sc_compute_window_naf(a, width):
    out = np.zeros(256, dtype=np.int8)
    lib().zo_sc_compute_window_naf(_p(_a(a, 5)), ctypes.c_uint(width), out.ctypes.data_as(_i8p))
    return out


# ---- EdwardsPoint / RistrettoPoint -------------------------------------------------------------------
def pt_identity():
    out = np.zeros(20, dtype=np.uint64)
    lib().zo_pt_identity(_p(out))
    return out


def pt_neg(p): return _un("zo_pt_neg", p, 20, 20)
def pt_double(p): return _un("zo_pt_double", p, 20, 20)
def pt_add(p, q): return _bin("zo_pt_add", p, q, 20, 20, 20)
def pt_sub(p, q): return _bin("zo_pt_sub", p, q, 20, 20, 20)
def pt_double_and_add(p, s): return _bin("zo_pt_double_and_add", p, s, 20, 5, 20)
def pt_ltr_bin_mul(p, s): return _bin("zo_pt_ltr_bin_mul", p, s, 20, 5, 20)
def pt_binary_naf_mul(p, s): return _bin("zo_pt_binary_naf_mul", p, s, 20, 5, 20)
def pt_eq(p, q): return int(lib().zo_pt_eq(_p(_a(p, 20)), _p(_a(q, 20))))
def pt_is_valid(p): return int(lib().zo_pt_is_valid(_p(_a(p, 20))))
def ris_eq(p, q): return int(lib().zo_ris_eq(_p(_a(p, 20)), _p(_a(q, 20))))


def pt_to_affine(p):
    out = np.zeros(10, dtype=np.uint64)
    if lib().zo_pt_to_affine(_p(_a(p, 20)), _p(out)):
        raise ZeroDivisionError("Z == 0")
    return out


def pt_new_from_y_coord(y, sign):
    out = np.zeros(20, dtype=np.uint64)
    ok = lib().zo_pt_new_from_y_coord(_p(_a(y, 5)), int(sign), _p(out))
    return out if ok else None


def pt_compress(p):
    out = np.zeros(32, dtype=np.uint8)
    if lib().zo_pt_compress(_p(_a(p, 20)), out.ctypes.data_as(_u8p)):
        raise ValueError("compress failed")
    return out.tobytes()


def pt_decompress(b):
    b = _bytes_in(b)
    out = np.zeros(20, dtype=np.uint64)
    ok = lib().zo_pt_decompress(b.ctypes.data_as(_u8p), _p(out))
    return out if ok else None


def ris_compress(p):
    out = np.zeros(32, dtype=np.uint8)
    lib().zo_ris_compress(_p(_a(p, 20)), out.ctypes.data_as(_u8p))
    return out.tobytes()


def ris_decompress(b):
    b = _bytes_in(b)
    out = np.zeros(20, dtype=np.uint64)
    ok = lib().zo_ris_decompress(b.ctypes.data_as(_u8p), _p(out))
    return out if ok else None


def ris_elligator(r0):
    return _un("zo_ris_elligator", r0, 5, 20)


def ris_from_uniform_bytes(b):
    b = _bytes_in(b)
    assert b.size == 64
    out = np.zeros(20, dtype=np.uint64)
    lib().zo_ris_from_uniform_bytes(b.ctypes.data_as(_u8p), _p(out))
    return out


# ---- batch drivers ---------------------------------------------------------------------------------
def _batch(name, ins, cols_out, threads=1):
    ins = [_a(x) for x in ins]
    n = ins[0].shape[0]
    out = np.zeros((n, cols_out), dtype=np.uint64)
    getattr(lib(), name)(*[_p(x) for x in ins], _p(out), ctypes.c_size_t(n), ctypes.c_int(threads))
    return out


def fe_mul_batch(a, b, threads=1): return _batch("zo_fe_mul_batch", [a, b], 5, threads)
def fe_square_batch(a, threads=1): return _batch("zo_fe_square_batch", [a], 5, threads)
def fe_add_batch(a, b, threads=1): return _batch("zo_fe_add_batch", [a, b], 5, threads)
def fe_sub_batch(a, b, threads=1): return _batch("zo_fe_sub_batch", [a, b], 5, threads)
def fe_neg_batch(a, threads=1): return _batch("zo_fe_neg_batch", [a], 5, threads)
def sc_mul_batch(a, b, threads=1): return _batch("zo_sc_mul_batch", [a, b], 5, threads)
def sc_square_batch(a, threads=1): return _batch("zo_sc_square_batch", [a], 5, threads)
def sc_add_batch(a, b, threads=1): return _batch("zo_sc_add_batch", [a, b], 5, threads)
def sc_sub_batch(a, b, threads=1): return _batch("zo_sc_sub_batch", [a, b], 5, threads)
def pt_add_batch(p, q, threads=1): return _batch("zo_pt_add_batch", [p, q], 20, threads)
def pt_sub_batch(p, q, threads=1): return _batch("zo_pt_sub_batch", [p, q], 20, threads)
def pt_double_batch(p, threads=1): return _batch("zo_pt_double_batch", [p], 20, threads)
def pt_neg_batch(p, threads=1): return _batch("zo_pt_neg_batch", [p], 20, threads)
def pt_scalar_mul_batch(p, s, threads=1): return _batch("zo_pt_scalar_mul_batch", [p, s], 20, threads)
def pt_to_affine_batch(p, threads=1): return _batch("zo_pt_to_affine_batch", [p], 10, threads)


def fe_mul_square_batch(a, b, threads=1, out=None):
    """out = (prod, sq): caller-owned, already touched (n, 5) uint64 arrays -- keeps page faults out of a timed call."""
    a, b = _a(a, 5), _a(b, 5)
    n = a.shape[0]
    if out is None:
        prod = np.zeros((n, 5), dtype=np.uint64)
        sq = np.zeros((n, 5), dtype=np.uint64)
    else:
        prod, sq = out
        assert prod.shape == (n, 5) and sq.shape == (n, 5) and prod.dtype == np.uint64 and sq.dtype == np.uint64
        assert prod.flags.c_contiguous and sq.flags.c_contiguous
    lib().zo_fe_mul_square_batch(_p(a), _p(b), _p(prod), _p(sq), ctypes.c_size_t(n), ctypes.c_int(threads))
    return prod, sq


def ris_compress_batch(p, threads=1):
    p = _a(p, 20)
    n = p.shape[0]
    out = np.zeros((n, 32), dtype=np.uint8)
    lib().zo_ris_compress_batch(_p(p), out.ctypes.data_as(_u8p), ctypes.c_size_t(n), ctypes.c_int(threads))
    return out


def msm_naive(p, s, threads=1):
    p, s = _a(p, 20), _a(s, 5)
    n = p.shape[0]
    out = np.zeros(20, dtype=np.uint64)
    lib().zo_msm_naive(_p(p), _p(s), ctypes.c_size_t(n), ctypes.c_int(threads), _p(out))
    return out


def synth_fe(seed, stream, first, n):
    out = np.zeros((n, 5), dtype=np.uint64)
    lib().zo_synth_fe(ctypes.c_uint64(seed), ctypes.c_uint64(stream), ctypes.c_size_t(first), ctypes.c_size_t(n), _p(out))
    return out


def synth_scalar(seed, stream, first, n):
    out = np.zeros((n, 5), dtype=np.uint64)
    lib().zo_synth_scalar(ctypes.c_uint64(seed), ctypes.c_uint64(stream), ctypes.c_size_t(first), ctypes.c_size_t(n), _p(out))
    return out


# ---- limb <-> int helpers for tests ------------------------------------------------------------------
def limbs_to_int(l):
    return sum(int(x) << (52 * i) for i, x in enumerate(np.asarray(l).reshape(-1)[:5]))


def int_to_limbs(v):
    return np.array([(v >> (52 * i)) & ((1 << 52) - 1) for i in range(5)], dtype=np.uint64)
