/*
 * zerocaf_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, unsigned __int128) of the reference crate's hot path, used as the
 * parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * leg.  Nothing under dusk_zerocaf_b200/ may include, link or load this.
 *
 * Parity is PINNED: tests/test_oracle_kats.py checks this library against every known-answer
 * vector the reference's own unit tests hold for the path (tests/golden/reference_kats.json,
 * extracted from /root/reference by tests/golden/extract_kats.py).
 *
 * Layout everywhere: FieldElement / Scalar = uint64_t[5], radix 2^52, little-endian limb order
 * (reference src/backend/u64/field.rs:31-32, scalar.rs:26-27); EdwardsPoint = uint64_t[20] =
 * X|Y|Z|T (reference src/edwards.rs:336-342).
 */
#ifndef ZEROCAF_ORACLE_H
#define ZEROCAF_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t v[5]; } zo_fe;     /* FieldElement([u64;5]) */
typedef struct { uint64_t v[5]; } zo_sc;     /* Scalar([u64;5])       */
typedef struct { zo_fe X, Y, Z, T; } zo_pt;  /* EdwardsPoint          */

/* ---- FieldElement (src/backend/u64/field.rs) ---- */
void zo_fe_add(const uint64_t a[5], const uint64_t b[5], uint64_t out[5]);
void zo_fe_sub(const uint64_t a[5], const uint64_t b[5], uint64_t out[5]);
void zo_fe_neg(const uint64_t a[5], uint64_t out[5]);
void zo_fe_mul(const uint64_t a[5], const uint64_t b[5], uint64_t out[5]);
void zo_fe_square(const uint64_t a[5], uint64_t out[5]);
void zo_fe_montgomery_mul(const uint64_t a[5], const uint64_t b[5], uint64_t out[5]);
void zo_fe_to_montgomery(const uint64_t a[5], uint64_t out[5]);
void zo_fe_from_montgomery(const uint64_t a[5], uint64_t out[5]);
void zo_fe_from_bytes(const uint8_t bytes[32], uint64_t out[5]);
void zo_fe_to_bytes(const uint64_t a[5], uint8_t out[32]);
void zo_fe_half_without_mod(const uint64_t a[5], uint64_t out[5]);
void zo_fe_half(const uint64_t a[5], uint64_t out[5]);
void zo_fe_two_pow_k(uint64_t k, uint64_t out[5]);
int  zo_fe_inverse(const uint64_t a[5], uint64_t out[5]);                 /* 0 ok, 1 = inverse of zero (reference panics) */
int  zo_fe_div(const uint64_t a[5], const uint64_t b[5], uint64_t out[5]); /* 1 = division by zero */
void zo_fe_pow(const uint64_t a[5], const uint64_t e[5], uint64_t out[5]);
int  zo_fe_legendre_symbol(const uint64_t a[5]);                          /* Choice as 0/1 */
int  zo_fe_mod_sqrt(const uint64_t a[5], int sign, uint64_t out[5]);      /* 1 = Some, 0 = None */
int  zo_fe_sqrt_ratio_i(const uint64_t u[5], const uint64_t v[5], uint64_t out[5]); /* returns Choice */
int  zo_fe_inv_sqrt(const uint64_t a[5], uint64_t out[5]);
int  zo_fe_is_positive(const uint64_t a[5]);
int  zo_fe_cmp(const uint64_t a[5], const uint64_t b[5]);

/* ---- Scalar (src/backend/u64/scalar.rs) ---- */
void zo_sc_add(const uint64_t a[5], const uint64_t b[5], uint64_t out[5]);
void zo_sc_sub(const uint64_t a[5], const uint64_t b[5], uint64_t out[5]);
void zo_sc_neg(const uint64_t a[5], uint64_t out[5]);
void zo_sc_mul(const uint64_t a[5], const uint64_t b[5], uint64_t out[5]);
void zo_sc_square(const uint64_t a[5], uint64_t out[5]);
void zo_sc_montgomery_mul(const uint64_t a[5], const uint64_t b[5], uint64_t out[5]);
void zo_sc_to_montgomery(const uint64_t a[5], uint64_t out[5]);
void zo_sc_from_montgomery(const uint64_t a[5], uint64_t out[5]);
int  zo_sc_from_bytes(const uint8_t bytes[32], uint64_t out[5]);          /* 1 = value > L-1 (reference asserts) */
void zo_sc_to_bytes(const uint64_t a[5], uint8_t out[32]);
void zo_sc_half_without_mod(const uint64_t a[5], uint64_t out[5]);
void zo_sc_half(const uint64_t a[5], uint64_t out[5]);
void zo_sc_shr(const uint64_t a[5], unsigned n, uint64_t out[5]);
void zo_sc_pow(const uint64_t a[5], const uint64_t e[5], uint64_t out[5]);
void zo_sc_two_pow_k(uint64_t k, uint64_t out[5]);
void zo_sc_from_i8(int8_t x, uint64_t out[5]);
void zo_sc_into_bits(const uint64_t a[5], uint8_t bits[256]);
void zo_sc_compute_naf(const uint64_t a[5], int8_t naf[256]);
void zo_sc_compute_window_naf(const uint64_t a[5], unsigned width, int8_t naf[256]);

/* ---- EdwardsPoint (src/edwards.rs) ---- */
void zo_pt_identity(uint64_t out[20]);
void zo_pt_neg(const uint64_t p[20], uint64_t out[20]);
void zo_pt_add(const uint64_t p[20], const uint64_t q[20], uint64_t out[20]);
void zo_pt_sub(const uint64_t p[20], const uint64_t q[20], uint64_t out[20]);
void zo_pt_double(const uint64_t p[20], uint64_t out[20]);
void zo_pt_double_and_add(const uint64_t p[20], const uint64_t s[5], uint64_t out[20]);
void zo_pt_ltr_bin_mul(const uint64_t p[20], const uint64_t s[5], uint64_t out[20]);
void zo_pt_binary_naf_mul(const uint64_t p[20], const uint64_t s[5], uint64_t out[20]);
int  zo_pt_to_affine(const uint64_t p[20], uint64_t xy[10]);              /* 1 = Z == 0 */
int  zo_pt_eq(const uint64_t p[20], const uint64_t q[20]);                /* affine equality, edwards.rs:360-364 */
int  zo_pt_is_valid(const uint64_t p[20]);                                /* via ProjectivePoint, edwards.rs:393-400,733-748 */
int  zo_pt_new_from_y_coord(const uint64_t y[5], int sign, uint64_t out[20]);
int  zo_pt_compress(const uint64_t p[20], uint8_t out[32]);               /* CompressedEdwardsY */
int  zo_pt_decompress(const uint8_t in[32], uint64_t out[20]);

/* ---- RistrettoPoint (src/ristretto.rs) ---- */
int  zo_ris_eq(const uint64_t p[20], const uint64_t q[20]);               /* ristretto.rs:166-176 */
void zo_ris_compress(const uint64_t p[20], uint8_t out[32]);              /* ristretto.rs:398-425 */
int  zo_ris_decompress(const uint8_t in[32], uint64_t out[20]);           /* ristretto.rs:96-154 */
void zo_ris_elligator(const uint64_t r0[5], uint64_t out[20]);            /* ristretto.rs:430-471 */
void zo_ris_from_uniform_bytes(const uint8_t in[64], uint64_t out[20]);   /* ristretto.rs:493-507 */

/* ---- batch drivers (index-parallel over `threads` pthreads; threads<=1 -> serial) ----
 * These are what the CPU baseline times; semantics = the per-element functions above in a loop. */
void zo_fe_mul_batch(const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n, int threads);
void zo_fe_square_batch(const uint64_t *a, uint64_t *out, size_t n, int threads);
void zo_fe_add_batch(const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n, int threads);
void zo_fe_sub_batch(const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n, int threads);
void zo_fe_neg_batch(const uint64_t *a, uint64_t *out, size_t n, int threads);
void zo_fe_mul_square_batch(const uint64_t *a, const uint64_t *b, uint64_t *prod, uint64_t *sq, size_t n, int threads);
void zo_sc_mul_batch(const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n, int threads);
void zo_sc_square_batch(const uint64_t *a, uint64_t *out, size_t n, int threads);
void zo_sc_add_batch(const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n, int threads);
void zo_sc_sub_batch(const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n, int threads);
void zo_pt_add_batch(const uint64_t *p, const uint64_t *q, uint64_t *out, size_t n, int threads);
void zo_pt_sub_batch(const uint64_t *p, const uint64_t *q, uint64_t *out, size_t n, int threads);
void zo_pt_double_batch(const uint64_t *p, uint64_t *out, size_t n, int threads);
void zo_pt_neg_batch(const uint64_t *p, uint64_t *out, size_t n, int threads);
void zo_pt_scalar_mul_batch(const uint64_t *p, const uint64_t *s, uint64_t *out, size_t n, int threads);
void zo_pt_to_affine_batch(const uint64_t *p, uint64_t *xy, size_t n, int threads);
void zo_ris_compress_batch(const uint64_t *p, uint8_t *out, size_t n, int threads);
/* MSM has no reference implementation; the derived oracle is
 * fold(Add, identity, [double_and_add(P_i, s_i)]) in index order. With threads > 1
 * each thread folds a contiguous index range and the partials are folded in thread order; the result
 * is then equal to the serial one as a group element (compare affine / compressed), not limb-wise. */
void zo_msm_naive(const uint64_t *p, const uint64_t *s, size_t n, int threads, uint64_t out[20]);

/* ---- deterministic synthetic inputs (SplitMix64 counter PRNG) ---- */
void zo_synth_fe(uint64_t seed, uint64_t stream, size_t first, size_t n, uint64_t *out);      /* FieldElement::random  src/field.rs:131-140 */
void zo_synth_scalar(uint64_t seed, uint64_t stream, size_t first, size_t n, uint64_t *out);  /* Scalar::random        src/scalar.rs:100-109 */

#ifdef __cplusplus
}
#endif
#endif
