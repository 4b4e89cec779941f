/*
 * zerocaf_oracle.c -- TEST INFRASTRUCTURE ONLY (see zerocaf_oracle.h).
 *
 * A CPU restatement of the reference's algorithms for the FieldElement / Scalar / EdwardsPoint /
 * RistrettoPoint hot path, following the reference SCHEDULE one-to-one (radix-2^52 limbs, 5x5
 * schoolbook into nine u128 columns, Montgomery reduction with R = 2^260 executed twice per
 * multiplication, add-as-double, LSB-first double-and-add), so that
 *   (1) every value it returns is limb-identical to what the Rust crate returns, and
 *   (2) timing it is a fair stand-in for the crate's CPU speed (the Rust toolchain is not in this
 *       image; see DESIGN.md "Oracle").
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 *
 * Build: see oracle/Makefile (gcc -O3 -march=native -flto, mirroring Cargo.toml:47-54).
 */
#include "zerocaf_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;

#define MASK52 ((((u64)1) << 52) - 1)

/* ------------------------------------------------------------------------------------------- */
/* Constants: src/backend/u64/constants.rs                                                      */
/* ------------------------------------------------------------------------------------------- */
static const zo_fe FIELD_L = {{671914833335277ULL, 3916664325105025ULL, 1367801ULL, 0ULL, 17592186044416ULL}};            /* :30-36 */
static const u64 LFACTOR_FIELD = 1439961107955227ULL;                                                                      /* :59    */
static const zo_fe RR_FIELD = {{2764609938444603ULL, 3768881411696287ULL, 1616719297148420ULL, 1087343033131391ULL, 10175238647962ULL}}; /* :39-45 */
static const zo_fe INVERSE_MOD_TWO = {{2587757230352887ULL, 4210131976237760ULL, 683900ULL, 0ULL, 8796093022208ULL}};      /* :52    */
static const zo_fe MINUS_ONE_HALF = {{2587757230352886ULL, 4210131976237760ULL, 683900ULL, 0ULL, 8796093022208ULL}};       /* :55    */
static const zo_fe POS_RANGE = {{2587757230352886ULL, 4210131976237760ULL, 683900ULL, 0ULL, 8796093022208ULL}};            /* :12-13 */
static const zo_fe EDWARDS_A = {{671914833335276ULL, 3916664325105025ULL, 1367801ULL, 0ULL, 17592186044416ULL}};           /* :75-81 */
static const zo_fe EDWARDS_D = {{3304133203739795ULL, 2446467598308289ULL, 1534112949566882ULL, 2032729967918914ULL, 2313225441931ULL}}; /* :86-92 */
static const zo_fe SQRT_MINUS_ONE = {{3075585030474777ULL, 2451921961843096ULL, 1194333869305507ULL, 2218299809671669ULL, 7376823328646ULL}}; /* :95-101 */
static const zo_fe INV_SQRT_A_MINUS_D = {{550050132044477ULL, 3953042081665262ULL, 2971403105229349ULL, 212915494370164ULL, 1172367057772ULL}}; /* :122-128 */
static const zo_fe SQRT_AD_MINUS_ONE = {{3601277882726560ULL, 1817821323014817ULL, 1726005090908779ULL, 2111284621343800ULL, 648674458156ULL}}; /* :131-137 */

static const zo_sc SC_L = {{1129677152307299ULL, 1363544697812651ULL, 714439ULL, 0ULL, 2199023255552ULL}};                 /* :9     */
static const u64 LFACTOR = 1331240223835829ULL;                                                                            /* :18    */
static const zo_sc SC_RR = {{137682194168839ULL, 3209056245311277ULL, 1480926248458276ULL, 2533620989757837ULL, 1314911199310ULL}}; /* :21-27 */
static const zo_sc SCALAR_INVERSE_MOD_TWO = {{2816638389838898ULL, 2933572162591573ULL, 357219ULL, 0ULL, 1099511627776ULL}}; /* :48 */
static const zo_sc SC_MINUS_ONE = {{1129677152307298ULL, 1363544697812651ULL, 714439ULL, 0ULL, 2199023255552ULL}};         /* scalar.rs:341-343 */

static const zo_fe FE_ZERO = {{0, 0, 0, 0, 0}};
static const zo_fe FE_ONE = {{1, 0, 0, 0, 0}};
static const zo_fe FE_MINUS_ONE = {{671914833335276ULL, 3916664325105025ULL, 1367801ULL, 0ULL, 17592186044416ULL}};        /* field.rs:524-532 */
static const zo_sc SC_ZERO = {{0, 0, 0, 0, 0}};
static const zo_sc SC_ONE = {{1, 0, 0, 0, 0}};

static inline u128 m(u64 x, u64 y) { return (u128)x * (u128)y; } /* field.rs:506-508, scalar.rs:325-327 */

/* ------------------------------------------------------------------------------------------- */
/* Shared limb helpers. FieldElement and Scalar use the same radix and the same add/sub/reduce  */
/* skeleton with a different modulus (field.rs:191-240 vs scalar.rs:184-237).                   */
/* ------------------------------------------------------------------------------------------- */

/* Ord::cmp, most significant limb first: field.rs:66-77, scalar.rs:53-64 */
static int limbs_cmp(const u64 *a, const u64 *b) {
    for (int i = 4; i >= 0; i--) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return -1;
    }
    return 0;
}

/* a - b (mod modulus): wrapping limb subtraction with the borrow in bit 63, then a masked add of
 * the modulus when the last limb underflowed.  field.rs:217-240 / scalar.rs:210-237 */
static void limbs_sub(const u64 *a, const u64 *b, const u64 *modulus, u64 *out) {
    u64 sub = 0, diff[5];
    for (int i = 0; i < 5; i++) {
        sub = a[i] - (b[i] + (sub >> 63));
        diff[i] = sub & MASK52;
    }
    u64 underflow_mask = ((sub >> 63) ^ 1) - 1;
    u64 carry = 0;
    for (int i = 0; i < 5; i++) {
        carry = (carry >> 52) + diff[i] + (modulus[i] & underflow_mask);
        out[i] = carry & MASK52;
    }
}

/* a + b (mod modulus): limb add with 52-bit carries, then `sum - modulus` through limbs_sub.
 * field.rs:191-207 / scalar.rs:184-200 */
static void limbs_add(const u64 *a, const u64 *b, const u64 *modulus, u64 *out) {
    u64 sum[5], carry = 0;
    for (int i = 0; i < 5; i++) {
        carry = a[i] + b[i] + (carry >> 52);
        sum[i] = carry & MASK52;
    }
    limbs_sub(sum, modulus, modulus, out);
}

/* 5x5 schoolbook product into nine un-carried u128 columns. field.rs:741-757 / scalar.rs:580-594 */
static void limbs_mul_internal(const u64 *a, const u64 *b, u128 r[9]) {
    r[0] = m(a[0], b[0]);
    r[1] = m(a[0], b[1]) + m(a[1], b[0]);
    r[2] = m(a[0], b[2]) + m(a[1], b[1]) + m(a[2], b[0]);
    r[3] = m(a[0], b[3]) + m(a[1], b[2]) + m(a[2], b[1]) + m(a[3], b[0]);
    r[4] = m(a[0], b[4]) + m(a[1], b[3]) + m(a[2], b[2]) + m(a[3], b[1]) + m(a[4], b[0]);
    r[5] = m(a[1], b[4]) + m(a[2], b[3]) + m(a[3], b[2]) + m(a[4], b[1]);
    r[6] = m(a[2], b[4]) + m(a[3], b[3]) + m(a[4], b[2]);
    r[7] = m(a[3], b[4]) + m(a[4], b[3]);
    r[8] = m(a[4], b[4]);
}

/* Squaring with doubled low limbs. field.rs:763-777 / scalar.rs:600-614 */
static void limbs_square_internal(const u64 *a, u128 r[9]) {
    u64 d0 = a[0] * 2, d1 = a[1] * 2, d2 = a[2] * 2, d3 = a[3] * 2;
    r[0] = m(a[0], a[0]);
    r[1] = m(d0, a[1]);
    r[2] = m(d0, a[2]) + m(a[1], a[1]);
    r[3] = m(d0, a[3]) + m(d1, a[2]);
    r[4] = m(d0, a[4]) + m(d1, a[3]) + m(a[2], a[2]);
    r[5] = m(d1, a[4]) + m(d2, a[3]);
    r[6] = m(d2, a[4]) + m(a[3], a[3]);
    r[7] = m(d3, a[4]);
    r[8] = m(a[4], a[4]);
}

/* half_without_mod: shift right by one across 52-bit limbs. field.rs:676-688 / scalar.rs:562-574 */
static void limbs_half_without_mod(const u64 *a, u64 *out) {
    u64 carry = 0, res[5];
    memcpy(res, a, sizeof res);
    for (int i = 4; i >= 0; i--) {
        res[i] |= carry;
        carry = (res[i] & 1) << 52;
        res[i] >>= 1;
    }
    memcpy(out, res, sizeof res);
}

/* to_bytes: field.rs:591-631 / scalar.rs:477-516 (identical byte schedule) */
static void limbs_to_bytes(const u64 *s, uint8_t res[32]) {
    res[0] = (uint8_t)(s[0] >> 0);   res[1] = (uint8_t)(s[0] >> 8);   res[2] = (uint8_t)(s[0] >> 16);
    res[3] = (uint8_t)(s[0] >> 24);  res[4] = (uint8_t)(s[0] >> 32);  res[5] = (uint8_t)(s[0] >> 40);
    res[6] = (uint8_t)((s[0] >> 48) | (s[1] << 4));
    res[7] = (uint8_t)(s[1] >> 4);   res[8] = (uint8_t)(s[1] >> 12);  res[9] = (uint8_t)(s[1] >> 20);
    res[10] = (uint8_t)(s[1] >> 28); res[11] = (uint8_t)(s[1] >> 36); res[12] = (uint8_t)(s[1] >> 44);
    res[13] = (uint8_t)(s[2] >> 0);  res[14] = (uint8_t)(s[2] >> 8);  res[15] = (uint8_t)(s[2] >> 16);
    res[16] = (uint8_t)(s[2] >> 24); res[17] = (uint8_t)(s[2] >> 32); res[18] = (uint8_t)(s[2] >> 40);
    res[19] = (uint8_t)((s[2] >> 48) | (s[3] << 4));
    res[20] = (uint8_t)(s[3] >> 4);  res[21] = (uint8_t)(s[3] >> 12); res[22] = (uint8_t)(s[3] >> 20);
    res[23] = (uint8_t)(s[3] >> 28); res[24] = (uint8_t)(s[3] >> 36); res[25] = (uint8_t)(s[3] >> 44);
    res[26] = (uint8_t)(s[4] >> 0);  res[27] = (uint8_t)(s[4] >> 8);  res[28] = (uint8_t)(s[4] >> 16);
    res[29] = (uint8_t)(s[4] >> 24); res[30] = (uint8_t)(s[4] >> 32); res[31] = (uint8_t)(s[4] >> 40);
}

/* two_pow_k body shared by field.rs:637-666, :719-739 and scalar.rs:525-552 */
static void limbs_two_pow(u64 e, u64 *out) {
    memset(out, 0, 5 * sizeof(u64));
    if (e <= 51) out[0] = (u64)1 << e;
    else if (e <= 103) out[1] = (u64)1 << (e - 52);
    else if (e <= 155) out[2] = (u64)1 << (e - 104);
    else if (e <= 207) out[3] = (u64)1 << (e - 156);
    else out[4] = (u64)1 << (e - 208);
}

/* ------------------------------------------------------------------------------------------- */
/* FieldElement: src/backend/u64/field.rs                                                       */
/* ------------------------------------------------------------------------------------------- */
static inline zo_fe fe_add(const zo_fe *a, const zo_fe *b) { zo_fe r; limbs_add(a->v, b->v, FIELD_L.v, r.v); return r; } /* :191-207 */
static inline zo_fe fe_sub(const zo_fe *a, const zo_fe *b) { zo_fe r; limbs_sub(a->v, b->v, FIELD_L.v, r.v); return r; } /* :217-240 */
static inline zo_fe fe_neg(const zo_fe *a) { return fe_sub(&FE_ZERO, a); }                                                 /* :170-179 */

/* limbs / R (mod p), R = 2^260; terms with FIELD_L[3] == 0 are skipped exactly as the reference
 * does. field.rs:780-813 */
static zo_fe fe_montgomery_reduce(const u128 limbs[9]) {
    const u64 *l = FIELD_L.v;
    u128 carry, sum;
    u64 n0, n1, n2, n3, n4, r[5];
#define ADJ(sumexpr, n)                                                  \
    sum = (sumexpr);                                                     \
    n = ((u64)sum * LFACTOR_FIELD) & MASK52;                              \
    carry = (sum + m(n, l[0])) >> 52;
#define RES(sumexpr, w)                                                  \
    sum = (sumexpr);                                                     \
    w = (u64)sum & MASK52;                                               \
    carry = sum >> 52;
    ADJ(limbs[0], n0)
    ADJ(carry + limbs[1] + m(n0, l[1]), n1)
    ADJ(carry + limbs[2] + m(n0, l[2]) + m(n1, l[1]), n2)
    ADJ(carry + limbs[3] + m(n1, l[2]) + m(n2, l[1]), n3)
    ADJ(carry + limbs[4] + m(n0, l[4]) + m(n2, l[2]) + m(n3, l[1]), n4)
    RES(carry + limbs[5] + m(n1, l[4]) + m(n3, l[2]) + m(n4, l[1]), r[0])
    RES(carry + limbs[6] + m(n2, l[4]) + m(n4, l[2]), r[1])
    RES(carry + limbs[7] + m(n3, l[4]), r[2])
    RES(carry + limbs[8] + m(n4, l[4]), r[3])
    r[4] = (u64)carry;
#undef ADJ
#undef RES
    zo_fe out;
    limbs_sub(r, l, l, out.v);
    return out;
}

static inline zo_fe fe_montgomery_mul(const zo_fe *a, const zo_fe *b) { /* :818-820 */
    u128 t[9];
    limbs_mul_internal(a->v, b->v, t);
    return fe_montgomery_reduce(t);
}

/* Mul: mred(mred(a*b) * RR_FIELD). field.rs:250-262 */
static inline zo_fe fe_mul(const zo_fe *a, const zo_fe *b) {
    zo_fe prod = fe_montgomery_mul(a, b);
    return fe_montgomery_mul(&prod, &RR_FIELD);
}

/* Square: mred(mred(a^2) * RR_FIELD). field.rs:302-315 */
static inline zo_fe fe_square(const zo_fe *a) {
    u128 t[9];
    limbs_square_internal(a->v, t);
    zo_fe aa = fe_montgomery_reduce(t);
    return fe_montgomery_mul(&aa, &RR_FIELD);
}

static inline zo_fe fe_to_montgomery(const zo_fe *a) { return fe_montgomery_mul(a, &RR_FIELD); } /* :824-826 */
static inline zo_fe fe_from_montgomery(const zo_fe *a) {                                          /* :830-836 */
    u128 t[9] = {0};
    for (int i = 0; i < 5; i++) t[i] = a->v[i];
    return fe_montgomery_reduce(t);
}

static inline int fe_is_even(const zo_fe *a) { return (a->v[0] & 1) == 0; }                      /* :534-539 */
static inline zo_fe fe_half_without_mod(const zo_fe *a) { zo_fe r; limbs_half_without_mod(a->v, r.v); return r; }
static inline zo_fe fe_half(const zo_fe *a) { return fe_mul(a, &INVERSE_MOD_TWO); }              /* :317-323 */
static inline int fe_cmp(const zo_fe *a, const zo_fe *b) { return limbs_cmp(a->v, b->v); }

/* equality goes through to_bytes (src/field.rs:93-106) */
static int fe_eq(const zo_fe *a, const zo_fe *b) {
    uint8_t x[32], y[32];
    limbs_to_bytes(a->v, x);
    limbs_to_bytes(b->v, y);
    return memcmp(x, y, 32) == 0;
}

/* from_bytes: five overlapping 8-byte little-endian loads. field.rs:563-587 */
static zo_fe fe_from_bytes(const uint8_t *bytes) {
    u64 w[5];
    static const int off[5] = {0, 6, 12, 19, 24};
    static const int sh[5] = {0, 4, 8, 4, 16};
    zo_fe r;
    for (int k = 0; k < 5; k++) {
        w[k] = 0;
        for (int j = 0; j < 8; j++) w[k] |= (u64)bytes[off[k] + j] << (8 * j);
        r.v[k] = (w[k] >> sh[k]) & MASK52;
    }
    return r;
}

/* is_positive: value in [0, POS_RANGE]. field.rs:552-557 */
static int fe_is_positive(const zo_fe *a) {
    return fe_cmp(a, &FE_ZERO) >= 0 && fe_cmp(a, &POS_RANGE) <= 0;
}

/* Savas-Koc almost-Montgomery inverse. field.rs:854-925 */
static int fe_inverse(const zo_fe *a, zo_fe *out) {
    if (fe_eq(a, &FE_ZERO)) return 1; /* assert!(a != zero), :864 */
    zo_fe p = FIELD_L, u = FIELD_L, v = *a, r = FE_ZERO, s = FE_ONE;
    const zo_fe two = {{2, 0, 0, 0, 0}};
    u64 k = 0;
    while (fe_cmp(&v, &FE_ZERO) > 0) {
        if (fe_is_even(&u)) {
            u = fe_half_without_mod(&u);
            s = fe_mul(&s, &two);
        } else if (fe_is_even(&v)) {
            v = fe_half_without_mod(&v);
            r = fe_mul(&r, &two);
        } else if (fe_cmp(&u, &v) > 0) {
            u = fe_sub(&u, &v);
            u = fe_half_without_mod(&u);
            r = fe_add(&r, &s);
            s = fe_mul(&s, &two);
        } else { /* v >= u */
            v = fe_sub(&v, &u);
            v = fe_half_without_mod(&v);
            s = fe_add(&r, &s);
            r = fe_mul(&r, &two);
        }
        k += 1;
    }
    if (fe_cmp(&r, &p) > 0) r = fe_sub(&r, &p);
    r = fe_sub(&p, &r);
    /* phase 2: :917-924 */
    u64 z = k;
    if (z > 260) {
        r = fe_montgomery_mul(&r, &FE_ONE);
        z -= 260;
    }
    zo_fe fact;
    limbs_two_pow(260 - z, fact.v); /* inner_two_pow_k, :719-739 */
    *out = fe_montgomery_mul(&r, &fact);
    return 0;
}

/* Div: x * y^-1, asserts y != 0. field.rs:277-299 */
static int fe_div(const zo_fe *a, const zo_fe *b, zo_fe *out) {
    zo_fe inv;
    if (fe_inverse(b, &inv)) return 1;
    *out = fe_mul(a, &inv);
    return 0;
}

/* Pow: square-and-multiply driven by halving the exponent. field.rs:334-354 */
static zo_fe fe_pow(const zo_fe *a, const zo_fe *exp) {
    zo_fe base = *a, res = FE_ONE, e = *exp;
    while (fe_cmp(&e, &FE_ZERO) > 0) {
        if (fe_is_even(&e)) {
            e = fe_half_without_mod(&e);
            base = fe_mul(&base, &base);
        } else {
            e = fe_sub(&e, &FE_ONE);
            res = fe_mul(&res, &base);
            e = fe_half_without_mod(&e);
            base = fe_mul(&base, &base);
        }
    }
    return res;
}

/* legendre_symbol: a^((p-1)/2) != -1. field.rs:703-706 */
static int fe_legendre(const zo_fe *a) {
    zo_fe r = fe_pow(a, &MINUS_ONE_HALF);
    return fe_eq(&r, &FE_MINUS_ONE) ^ 1;
}

/* Tonelli-Shanks with the pre-computed non-residue 6. field.rs:378-440.
 * sign = 1 selects FIELD_L - x, sign = 0 selects x (conditional_select, :435-439). */
static int fe_mod_sqrt(const zo_fe *a, int sign, zo_fe *out) {
    if (fe_eq(a, &FE_ZERO)) { *out = FE_ZERO; return 1; }
    if (!fe_legendre(a)) return 0;
    const zo_fe one = FE_ONE, two = {{2, 0, 0, 0, 0}}, six = {{6, 0, 0, 0, 0}};
    zo_fe q = FE_MINUS_ONE, s = FE_ZERO;
    while (fe_is_even(&q)) {
        s = fe_add(&s, &one);
        q = fe_half_without_mod(&q);
    }
    zo_fe c = fe_pow(&six, &q);
    zo_fe q1 = fe_add(&q, &one);
    q1 = fe_half_without_mod(&q1);
    zo_fe x = fe_pow(a, &q1);
    zo_fe t = fe_pow(a, &q);
    zo_fe mm = s;
    while (!fe_eq(&t, &one)) {
        zo_fe i = FE_ZERO, e = {{2, 0, 0, 0, 0}};
        while (fe_cmp(&i, &mm) < 0) {
            i = fe_add(&i, &one);
            zo_fe te = fe_pow(&t, &e);
            if (fe_eq(&te, &one)) break;
            e = fe_mul(&e, &two);
        }
        zo_fe ex = fe_sub(&mm, &i);
        ex = fe_sub(&ex, &one);
        zo_fe tp = fe_pow(&two, &ex);
        zo_fe b = fe_pow(&c, &tp);
        x = fe_mul(&x, &b);
        zo_fe bb = fe_square(&b);
        t = fe_mul(&t, &bb);
        c = bb;
        mm = i;
    }
    if (sign) *out = fe_sub(&FIELD_L, &x);
    else *out = x;
    return 1;
}

/* sqrt_ratio_i: field.rs:474-502. Returns the Choice, writes the root. */
static int fe_sqrt_ratio_i(const zo_fe *u, const zo_fe *v, zo_fe *out) {
    int uz = fe_eq(u, &FE_ZERO), vz = fe_eq(v, &FE_ZERO);
    if (uz) { *out = FE_ZERO; return 1; }
    if (vz) { *out = FE_ZERO; return 0; }
    zo_fe ratio;
    fe_div(u, v, &ratio);
    zo_fe res;
    if (fe_legendre(&ratio)) {
        fe_mod_sqrt(&ratio, 1, &res);
        if (!fe_is_positive(&res)) res = fe_neg(&res);
        *out = res;
        return 1;
    }
    zo_fe ir = fe_mul(&SQRT_MINUS_ONE, &ratio);
    fe_mod_sqrt(&ir, 1, &res);
    if (!fe_is_positive(&res)) res = fe_neg(&res);
    *out = res;
    return 0;
}

static int fe_inv_sqrt(const zo_fe *a, zo_fe *out) { return fe_sqrt_ratio_i(&FE_ONE, a, out); } /* :443-459 */

/* ------------------------------------------------------------------------------------------- */
/* Scalar: src/backend/u64/scalar.rs                                                            */
/* ------------------------------------------------------------------------------------------- */
static inline zo_sc sc_add(const zo_sc *a, const zo_sc *b) { zo_sc r; limbs_add(a->v, b->v, SC_L.v, r.v); return r; }  /* :184-200 */
static inline zo_sc sc_sub(const zo_sc *a, const zo_sc *b) { zo_sc r; limbs_sub(a->v, b->v, SC_L.v, r.v); return r; }  /* :210-237 */
static inline zo_sc sc_neg(const zo_sc *a) { return sc_sub(&SC_ZERO, a); }                                             /* :137-147 */

/* scalar.rs:617-652 -- full 5x5 n*l, l[3] terms are NOT skipped here */
static zo_sc sc_montgomery_reduce(const u128 limbs[9]) {
    const u64 *l = SC_L.v;
    u128 carry, sum;
    u64 n0, n1, n2, n3, n4, r[5];
#define ADJ(sumexpr, n)                                                  \
    sum = (sumexpr);                                                     \
    n = ((u64)sum * LFACTOR) & MASK52;                                    \
    carry = (sum + m(n, l[0])) >> 52;
#define RES(sumexpr, w)                                                  \
    sum = (sumexpr);                                                     \
    w = (u64)sum & MASK52;                                               \
    carry = sum >> 52;
    ADJ(limbs[0], n0)
    ADJ(carry + limbs[1] + m(n0, l[1]), n1)
    ADJ(carry + limbs[2] + m(n0, l[2]) + m(n1, l[1]), n2)
    ADJ(carry + limbs[3] + m(n0, l[3]) + m(n1, l[2]) + m(n2, l[1]), n3)
    ADJ(carry + limbs[4] + m(n0, l[4]) + m(n1, l[3]) + m(n2, l[2]) + m(n3, l[1]), n4)
    RES(carry + limbs[5] + m(n1, l[4]) + m(n2, l[3]) + m(n3, l[2]) + m(n4, l[1]), r[0])
    RES(carry + limbs[6] + m(n2, l[4]) + m(n3, l[3]) + m(n4, l[2]), r[1])
    RES(carry + limbs[7] + m(n3, l[4]) + m(n4, l[3]), r[2])
    RES(carry + limbs[8] + m(n4, l[4]), r[3])
    r[4] = (u64)carry;
#undef ADJ
#undef RES
    zo_sc out;
    limbs_sub(r, l, l, out.v);
    return out;
}

static inline zo_sc sc_montgomery_mul(const zo_sc *a, const zo_sc *b) { /* :655-657 */
    u128 t[9];
    limbs_mul_internal(a->v, b->v, t);
    return sc_montgomery_reduce(t);
}
static inline zo_sc sc_mul(const zo_sc *a, const zo_sc *b) {            /* :247-258 */
    zo_sc ab = sc_montgomery_mul(a, b);
    return sc_montgomery_mul(&ab, &SC_RR);
}
static inline zo_sc sc_square(const zo_sc *a) {                         /* :272-283 */
    u128 t[9];
    limbs_square_internal(a->v, t);
    zo_sc aa = sc_montgomery_reduce(t);
    return sc_montgomery_mul(&aa, &SC_RR);
}
static inline zo_sc sc_to_montgomery(const zo_sc *a) { return sc_montgomery_mul(a, &SC_RR); } /* :661-663 */
static inline zo_sc sc_from_montgomery(const zo_sc *a) {                                       /* :667-673 */
    u128 t[9] = {0};
    for (int i = 0; i < 5; i++) t[i] = a->v[i];
    return sc_montgomery_reduce(t);
}
static inline int sc_is_even(const zo_sc *a) { return (a->v[0] & 1) == 0; }                   /* :346-348 */
static inline zo_sc sc_half_without_mod(const zo_sc *a) { zo_sc r; limbs_half_without_mod(a->v, r.v); return r; }
static inline zo_sc sc_half(const zo_sc *a) { return sc_mul(a, &SCALAR_INVERSE_MOD_TWO); }    /* :285-291 */
static inline int sc_cmp(const zo_sc *a, const zo_sc *b) { return limbs_cmp(a->v, b->v); }
static int sc_eq(const zo_sc *a, const zo_sc *b) { /* src/scalar.rs:78-91 */
    uint8_t x[32], y[32];
    limbs_to_bytes(a->v, x);
    limbs_to_bytes(b->v, y);
    return memcmp(x, y, 32) == 0;
}

/* from_bytes: four LE words re-cut into 52-bit limbs; value must be <= L-1. scalar.rs:445-467 */
static int sc_from_bytes(const uint8_t *bytes, zo_sc *out) {
    u64 w[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 8; j++) w[i] |= (u64)bytes[i * 8 + j] << (j * 8);
    const u64 top_mask = (((u64)1) << 48) - 1;
    zo_sc s;
    s.v[0] = w[0] & MASK52;
    s.v[1] = ((w[0] >> 52) | (w[1] << 12)) & MASK52;
    s.v[2] = ((w[1] >> 40) | (w[2] << 24)) & MASK52;
    s.v[3] = ((w[2] >> 28) | (w[3] << 36)) & MASK52;
    s.v[4] = (w[3] >> 16) & top_mask;
    *out = s;
    return sc_cmp(&s, &SC_MINUS_ONE) <= 0 ? 0 : 1;
}

/* From<i8>: scalar.rs:67-82 */
static zo_sc sc_from_i8(int8_t x) {
    zo_sc r = SC_ZERO;
    if (x >= 0) { r.v[0] = (u64)x; return r; }
    r.v[0] = (u64)(-(int)x);
    return sc_neg(&r);
}

/* Shr<u8>: repeated one-bit shifts. scalar.rs:161-181 */
static zo_sc sc_shr(const zo_sc *a, unsigned n) {
    zo_sc r = *a;
    for (unsigned k = 0; k < n; k++) r = sc_half_without_mod(&r);
    return r;
}

/* Pow: scalar.rs:293-322 (odd branch uses the modular half(), even branch half_without_mod) */
static zo_sc sc_pow(const zo_sc *a, const zo_sc *exp) {
    zo_sc base = *a, res = SC_ONE, e = *exp;
    while (sc_cmp(&e, &SC_ZERO) > 0) {
        if (sc_is_even(&e)) {
            e = sc_half_without_mod(&e);
            base = sc_square(&base);
        } else {
            e = sc_sub(&e, &SC_ONE);
            res = sc_mul(&res, &base);
            e = sc_half(&e);
            base = sc_square(&base);
        }
    }
    return res;
}

/* into_bits: LSB-first bits of to_bytes. scalar.rs:352-366 */
static void sc_into_bits(const zo_sc *a, uint8_t bits[256]) {
    uint8_t by[32];
    limbs_to_bytes(a->v, by);
    for (int k = 0; k < 32; k++)
        for (int i = 0; i < 8; i++) bits[8 * k + i] = (by[k] >> i) & 1;
}

static inline uint8_t sc_mod_2_pow_k(const zo_sc *a, unsigned k) { return (uint8_t)(a->v[0] & ((1u << k) - 1)); } /* :417-420 */
/* mods_2_pow_k: signed residue in [-2^(w-1), 2^(w-1)). scalar.rs:426-435 */
static int8_t sc_mods_2_pow_k(const zo_sc *a, unsigned w) {
    int8_t modulus = (int8_t)sc_mod_2_pow_k(a, w);
    int8_t half = (int8_t)(1 << (w - 1));
    if (modulus >= half) return (int8_t)(modulus - (int8_t)(uint8_t)(1u << w));
    return modulus;
}

/* compute_NAF: scalar.rs:370-390 */
static void sc_compute_naf(const zo_sc *a, int8_t naf[256]) {
    zo_sc k = *a;
    int i = 0;
    memset(naf, 0, 256);
    while (sc_cmp(&k, &SC_ONE) >= 0) {
        if (!sc_is_even(&k)) {
            int8_t ki = (int8_t)(2 - (int8_t)sc_mod_2_pow_k(&k, 2));
            naf[i] = ki;
            zo_sc d = sc_from_i8(ki);
            k = sc_sub(&k, &d);
        } else {
            naf[i] = 0;
        }
        k = sc_half_without_mod(&k);
        i++;
    }
}

/* compute_window_NAF: scalar.rs:396-415 */
static void sc_compute_window_naf(const zo_sc *a, unsigned width, int8_t naf[256]) {
    zo_sc k = *a;
    int i = 0;
    memset(naf, 0, 256);
    while (sc_cmp(&k, &SC_ONE) >= 0) {
        if (!sc_is_even(&k)) {
            int8_t ki = sc_mods_2_pow_k(&k, width);
            naf[i] = ki;
            zo_sc d = sc_from_i8(ki);
            k = sc_sub(&k, &d);
        } else {
            naf[i] = 0;
        }
        k = sc_half_without_mod(&k);
        i++;
    }
}

/* ------------------------------------------------------------------------------------------- */
/* EdwardsPoint: src/edwards.rs                                                                 */
/* ------------------------------------------------------------------------------------------- */
static const zo_pt PT_IDENTITY = {{{0, 0, 0, 0, 0}}, {{1, 0, 0, 0, 0}}, {{1, 0, 0, 0, 0}}, {{0, 0, 0, 0, 0}}}; /* :381-391 */

static zo_pt pt_neg(const zo_pt *p) { /* :440-455 */
    zo_pt r = *p;
    r.X = fe_neg(&p->X);
    r.T = fe_neg(&p->T);
    return r;
}

/* Unified extended addition, a = -1 (HWCD'08 sec. 3.1). edwards.rs:465-489 */
static zo_pt pt_add(const zo_pt *p, const zo_pt *q) {
    zo_fe A = fe_mul(&p->X, &q->X);
    zo_fe B = fe_mul(&p->Y, &q->Y);
    zo_fe dT = fe_mul(&EDWARDS_D, &p->T);
    zo_fe C = fe_mul(&dT, &q->T);
    zo_fe D = fe_mul(&p->Z, &q->Z);
    zo_fe s1 = fe_add(&p->X, &p->Y), s2 = fe_add(&q->X, &q->Y);
    zo_fe E = fe_mul(&s1, &s2);
    E = fe_sub(&E, &A);
    E = fe_sub(&E, &B);
    zo_fe F = fe_sub(&D, &C);
    zo_fe G = fe_add(&D, &C);
    zo_fe H = fe_add(&B, &A);
    zo_pt r;
    r.X = fe_mul(&E, &F);
    r.Y = fe_mul(&G, &H);
    r.Z = fe_mul(&F, &G);
    r.T = fe_mul(&E, &H);
    return r;
}

/* Sub: negate, same formulas, H = B - a*A. edwards.rs:503-531 */
static zo_pt pt_sub(const zo_pt *p, const zo_pt *q0) {
    zo_pt q = pt_neg(q0);
    zo_fe A = fe_mul(&p->X, &q.X);
    zo_fe B = fe_mul(&p->Y, &q.Y);
    zo_fe dT = fe_mul(&EDWARDS_D, &p->T);
    zo_fe C = fe_mul(&dT, &q.T);
    zo_fe D = fe_mul(&p->Z, &q.Z);
    zo_fe s1 = fe_add(&p->X, &p->Y), s2 = fe_add(&q.X, &q.Y);
    zo_fe E = fe_mul(&s1, &s2);
    E = fe_sub(&E, &A);
    E = fe_sub(&E, &B);
    zo_fe F = fe_sub(&D, &C);
    zo_fe G = fe_add(&D, &C);
    zo_fe aA = fe_mul(&EDWARDS_A, &A);
    zo_fe H = fe_sub(&B, &aA);
    zo_pt r;
    r.X = fe_mul(&E, &F);
    r.Y = fe_mul(&G, &H);
    r.Z = fe_mul(&F, &G);
    r.T = fe_mul(&E, &H);
    return r;
}

static inline zo_pt pt_double(const zo_pt *p) { return pt_add(p, p); } /* self + self, :589-591 */

/* double_and_add, LSB first; N is doubled once more after the top bit. edwards.rs:102-120 */
static zo_pt pt_double_and_add(const zo_pt *point, const zo_sc *scalar) {
    zo_pt N = *point, Q = PT_IDENTITY;
    zo_sc n = *scalar;
    while (!sc_eq(&n, &SC_ZERO)) {
        if (!sc_is_even(&n)) Q = pt_add(&Q, &N);
        N = pt_double(&N);
        n = sc_half_without_mod(&n);
    }
    return Q;
}

/* ltr_bin_mul: edwards.rs:122-134 (bits 248..0 only, as in the reference) */
static zo_pt pt_ltr_bin_mul(const zo_pt *point, const zo_sc *scalar) {
    uint8_t bits[256];
    sc_into_bits(scalar, bits);
    zo_pt Q = PT_IDENTITY;
    for (int i = 248; i >= 0; i--) {
        Q = pt_double(&Q);
        if (bits[i] == 1) Q = pt_add(&Q, point);
    }
    return Q;
}

/* binary_naf_mul: edwards.rs:136-153 */
static zo_pt pt_binary_naf_mul(const zo_pt *point, const zo_sc *scalar) {
    int8_t naf[256];
    sc_compute_naf(scalar, naf);
    zo_pt Q = PT_IDENTITY;
    for (int i = 249; i >= 0; i--) {
        Q = pt_double(&Q);
        if (naf[i] == 1) Q = pt_add(&Q, point);
        else if (naf[i] == -1) Q = pt_sub(&Q, point);
    }
    return Q;
}

/* AffinePoint::from(EdwardsPoint): edwards.rs:1085-1092 */
static int pt_to_affine(const zo_pt *p, zo_fe *x, zo_fe *y) {
    zo_fe zinv;
    if (fe_inverse(&p->Z, &zinv)) return 1;
    *x = fe_mul(&p->X, &zinv);
    *y = fe_mul(&p->Y, &zinv);
    return 0;
}

/* PartialEq via affine ct_eq: edwards.rs:360-371, 1044-1048 */
static int pt_eq(const zo_pt *p, const zo_pt *q) {
    zo_fe x1, y1, x2, y2;
    if (pt_to_affine(p, &x1, &y1) || pt_to_affine(q, &x2, &y2)) return -1;
    return fe_eq(&x1, &x2) && fe_eq(&y1, &y2);
}

/* is_valid through ProjectivePoint: (aX^2 + Y^2) Z^2 == Z^4 + d X^2 Y^2. edwards.rs:393-400, 733-748 */
static int pt_is_valid(const zo_pt *p) {
    zo_fe x2 = fe_square(&p->X), y2 = fe_square(&p->Y), z2 = fe_square(&p->Z);
    zo_fe ax2 = fe_mul(&EDWARDS_A, &x2);
    zo_fe l = fe_add(&ax2, &y2);
    l = fe_mul(&l, &z2);
    zo_fe z4 = fe_square(&z2);
    zo_fe dx2 = fe_mul(&EDWARDS_D, &x2);
    zo_fe r = fe_mul(&dx2, &y2);
    r = fe_add(&z4, &r);
    return fe_eq(&l, &r);
}

/* new_from_y_coord: ProjectivePoint (edwards.rs:949-967) lifted to extended (:402-413) */
static int pt_new_from_y_coord(const zo_fe *y, int sign, zo_pt *out) {
    zo_fe yy = fe_square(y);
    zo_fe num = fe_sub(&yy, &FE_ONE);
    zo_fe den = fe_mul(&EDWARDS_D, &yy);
    den = fe_sub(&den, &EDWARDS_A);
    zo_fe xx, x;
    if (fe_div(&num, &den, &xx)) return 0;
    if (!fe_mod_sqrt(&xx, sign, &x)) return 0;
    /* EdwardsPoint::from(ProjectivePoint{X:x, Y:y, Z:1}) */
    out->X = fe_mul(&x, &FE_ONE);
    out->Y = fe_mul(y, &FE_ONE);
    out->Z = fe_square(&FE_ONE);
    out->T = fe_mul(&x, y);
    return 1;
}

/* EdwardsPoint::compress -> CompressedEdwardsY. edwards.rs:613-629, find_xx :195-199 */
static int pt_compress(const zo_pt *p, uint8_t out[32]) {
    zo_fe x, y;
    if (pt_to_affine(p, &x, &y)) return 1;
    zo_fe yy = fe_square(&y);
    zo_fe a = fe_sub(&yy, &FE_ONE);
    zo_fe b = fe_mul(&EDWARDS_D, &yy);
    b = fe_sub(&b, &EDWARDS_A);
    zo_fe xx, res;
    if (fe_div(&a, &b, &xx)) return 1;
    if (!fe_mod_sqrt(&xx, 0, &res)) return 1;
    int sign = fe_eq(&res, &x) ? 0 : 1;
    limbs_to_bytes(y.v, out);
    out[31] |= (uint8_t)(sign << 7);
    return 0;
}

/* CompressedEdwardsY::decompress: edwards.rs:313-326 */
static int pt_decompress(const uint8_t in[32], zo_pt *out) {
    int sign = in[31] >> 7;
    uint8_t yb[32];
    memcpy(yb, in, 32);
    yb[31] &= 0x0f;
    zo_fe y = fe_from_bytes(yb);
    return pt_new_from_y_coord(&y, sign, out);
}

/* ------------------------------------------------------------------------------------------- */
/* RistrettoPoint: src/ristretto.rs                                                             */
/* ------------------------------------------------------------------------------------------- */

/* ct_eq: X1*Y2 == Y1*X2  or  X1*X2 == Y1*Y2. ristretto.rs:166-176 */
static int ris_eq(const zo_pt *p, const zo_pt *q) {
    zo_fe a1 = fe_mul(&p->X, &q->Y), a2 = fe_mul(&p->Y, &q->X);
    zo_fe b1 = fe_mul(&p->X, &q->X), b2 = fe_mul(&p->Y, &q->Y);
    return fe_eq(&a1, &a2) | fe_eq(&b1, &b2);
}

/* compress: ristretto.rs:398-425 */
static void ris_compress(const zo_pt *P, uint8_t out[32]) {
    zo_fe zy1 = fe_add(&P->Z, &P->Y), zy2 = fe_sub(&P->Z, &P->Y);
    zo_fe u1 = fe_mul(&zy1, &zy2);
    zo_fe u2 = fe_mul(&P->X, &P->Y);
    zo_fe u2sq = fe_square(&u2);
    zo_fe arg = fe_mul(&u1, &u2sq), I;
    fe_inv_sqrt(&arg, &I);
    zo_fe D1 = fe_mul(&u1, &I), D2 = fe_mul(&u2, &I);
    zo_fe Zinv = fe_mul(&D1, &D2);
    Zinv = fe_mul(&Zinv, &P->T);
    zo_fe tz = fe_mul(&P->T, &Zinv);
    zo_fe x, y, D;
    if (!fe_is_positive(&tz)) {
        x = fe_mul(&SQRT_MINUS_ONE, &P->Y);
        y = fe_mul(&SQRT_MINUS_ONE, &P->X);
        D = fe_mul(&D1, &INV_SQRT_A_MINUS_D);
    } else {
        x = P->X;
        y = P->Y;
        D = D2;
    }
    zo_fe xz = fe_mul(&x, &Zinv);
    if (!fe_is_positive(&xz)) y = fe_neg(&y);
    zo_fe s = fe_sub(&P->Z, &y);
    s = fe_mul(&s, &D);
    if (!fe_is_positive(&s)) s = fe_neg(&s);
    limbs_to_bytes(s.v, out);
}

/* decompress: ristretto.rs:96-154 */
static int ris_decompress(const uint8_t in[32], zo_pt *out) {
    zo_fe s = fe_from_bytes(in);
    uint8_t chk[32];
    limbs_to_bytes(s.v, chk);
    if (!fe_is_positive(&s) || memcmp(chk, in, 32) != 0) return 0;
    zo_fe ss = fe_square(&s);
    zo_fe u1 = fe_sub(&FE_ONE, &ss);
    zo_fe u2 = fe_add(&FE_ONE, &ss);
    zo_fe u2sq = fe_square(&u2);
    zo_fe u1sq = fe_square(&u1);
    zo_fe v = fe_mul(&EDWARDS_D, &u1sq);
    v = fe_neg(&v);
    v = fe_sub(&v, &u2sq);
    zo_fe arg = fe_mul(&v, &u2sq), I;
    if (!fe_inv_sqrt(&arg, &I)) return 0;
    zo_fe Dx = fe_mul(&I, &u2);
    zo_fe Dy = fe_mul(&I, &Dx);
    Dy = fe_mul(&Dy, &v);
    zo_fe s2 = fe_add(&s, &s);
    zo_fe x = fe_mul(&s2, &Dx);
    if (!fe_is_positive(&x)) x = fe_neg(&x);
    zo_fe y = fe_mul(&u1, &Dy);
    zo_fe t = fe_mul(&x, &y);
    if (!fe_is_positive(&t) || fe_eq(&y, &FE_ZERO)) return 0;
    out->X = x;
    out->Y = y;
    out->Z = FE_ONE;
    out->T = t;
    return 1;
}

/* elligator_ristretto_flavor: ristretto.rs:430-471 */
static zo_pt ris_elligator(const zo_fe *r0) {
    const zo_fe d = EDWARDS_D, one = FE_ONE;
    zo_fe c = fe_neg(&one);
    zo_fe dsq = fe_square(&d);
    zo_fe one_minus_d_sq = fe_sub(&one, &dsq);
    zo_fe r0sq = fe_square(r0);
    zo_fe r = fe_mul(&SQRT_MINUS_ONE, &r0sq);
    zo_fe rp1 = fe_add(&r, &one);
    zo_fe N_s = fe_mul(&rp1, &one_minus_d_sq);
    zo_fe dr = fe_mul(&d, &r);
    zo_fe t1 = fe_sub(&c, &dr);
    zo_fe t2 = fe_add(&r, &d);
    zo_fe D = fe_mul(&t1, &t2);
    zo_fe s;
    int is_sq = fe_sqrt_ratio_i(&N_s, &D, &s);
    zo_fe s_prim = fe_mul(&s, r0);
    if (fe_is_positive(&s_prim)) s_prim = fe_neg(&s_prim);
    if (!is_sq) { s = s_prim; c = r; }
    zo_fe rm1 = fe_sub(&r, &one);
    zo_fe dm1 = fe_sub(&d, &one);
    zo_fe dm1sq = fe_square(&dm1);
    zo_fe N_t = fe_mul(&c, &rm1);
    N_t = fe_mul(&N_t, &dm1sq);
    N_t = fe_sub(&N_t, &D);
    zo_fe ssq = fe_square(&s);
    zo_fe s2 = fe_add(&s, &s);
    zo_fe W0 = fe_mul(&s2, &D);
    zo_fe W1 = fe_mul(&N_t, &SQRT_AD_MINUS_ONE);
    zo_fe W2 = fe_sub(&one, &ssq);
    zo_fe W3 = fe_add(&one, &ssq);
    zo_pt P;
    P.X = fe_mul(&W0, &W3);
    P.Y = fe_mul(&W2, &W1);
    P.Z = fe_mul(&W1, &W3);
    P.T = fe_mul(&W0, &W2);
    return P;
}

/* from_uniform_bytes: ristretto.rs:493-507 */
static zo_pt ris_from_uniform_bytes(const uint8_t in[64]) {
    zo_fe r1 = fe_from_bytes(in), r2 = fe_from_bytes(in + 32);
    zo_pt R1 = ris_elligator(&r1), R2 = ris_elligator(&r2);
    return pt_add(&R1, &R2);
}

/* ------------------------------------------------------------------------------------------- */
/* extern "C" surface: thin array-in / array-out wrappers                                       */
/* ------------------------------------------------------------------------------------------- */
#define FE(p) ((const zo_fe *)(p))
#define SC(p) ((const zo_sc *)(p))
#define PT(p) ((const zo_pt *)(p))
static inline void put_fe(uint64_t *o, zo_fe r) { memcpy(o, r.v, 40); }
static inline void put_sc(uint64_t *o, zo_sc r) { memcpy(o, r.v, 40); }
static inline void put_pt(uint64_t *o, zo_pt r) { memcpy(o, &r, 160); }

void zo_fe_add(const uint64_t a[5], const uint64_t b[5], uint64_t out[5]) { put_fe(out, fe_add(FE(a), FE(b))); }
void zo_fe_sub(const uint64_t a[5], const uint64_t b[5], uint64_t out[5]) { put_fe(out, fe_sub(FE(a), FE(b))); }
void zo_fe_neg(const uint64_t a[5], uint64_t out[5]) { put_fe(out, fe_neg(FE(a))); }
void zo_fe_mul(const uint64_t a[5], const uint64_t b[5], uint64_t out[5]) { put_fe(out, fe_mul(FE(a), FE(b))); }
void zo_fe_square(const uint64_t a[5], uint64_t out[5]) { put_fe(out, fe_square(FE(a))); }
void zo_fe_montgomery_mul(const uint64_t a[5], const uint64_t b[5], uint64_t out[5]) { put_fe(out, fe_montgomery_mul(FE(a), FE(b))); }
void zo_fe_to_montgomery(const uint64_t a[5], uint64_t out[5]) { put_fe(out, fe_to_montgomery(FE(a))); }
void zo_fe_from_montgomery(const uint64_t a[5], uint64_t out[5]) { put_fe(out, fe_from_montgomery(FE(a))); }
void zo_fe_from_bytes(const uint8_t bytes[32], uint64_t out[5]) { put_fe(out, fe_from_bytes(bytes)); }
void zo_fe_to_bytes(const uint64_t a[5], uint8_t out[32]) { limbs_to_bytes(a, out); }
void zo_fe_half_without_mod(const uint64_t a[5], uint64_t out[5]) { put_fe(out, fe_half_without_mod(FE(a))); }
void zo_fe_half(const uint64_t a[5], uint64_t out[5]) { put_fe(out, fe_half(FE(a))); }
void zo_fe_two_pow_k(uint64_t k, uint64_t out[5]) { limbs_two_pow(k, out); }
int zo_fe_inverse(const uint64_t a[5], uint64_t out[5]) { zo_fe r; if (fe_inverse(FE(a), &r)) return 1; put_fe(out, r); return 0; }
int zo_fe_div(const uint64_t a[5], const uint64_t b[5], uint64_t out[5]) { zo_fe r; if (fe_div(FE(a), FE(b), &r)) return 1; put_fe(out, r); return 0; }
void zo_fe_pow(const uint64_t a[5], const uint64_t e[5], uint64_t out[5]) { put_fe(out, fe_pow(FE(a), FE(e))); }
int zo_fe_legendre_symbol(const uint64_t a[5]) { return fe_legendre(FE(a)); }
int zo_fe_mod_sqrt(const uint64_t a[5], int sign, uint64_t out[5]) { zo_fe r; if (!fe_mod_sqrt(FE(a), sign, &r)) return 0; put_fe(out, r); return 1; }
int zo_fe_sqrt_ratio_i(const uint64_t u[5], const uint64_t v[5], uint64_t out[5]) { zo_fe r; int c = fe_sqrt_ratio_i(FE(u), FE(v), &r); put_fe(out, r); return c; }
int zo_fe_inv_sqrt(const uint64_t a[5], uint64_t out[5]) { zo_fe r; int c = fe_inv_sqrt(FE(a), &r); put_fe(out, r); return c; }
int zo_fe_is_positive(const uint64_t a[5]) { return fe_is_positive(FE(a)); }
int zo_fe_cmp(const uint64_t a[5], const uint64_t b[5]) { return limbs_cmp(a, b); }

void zo_sc_add(const uint64_t a[5], const uint64_t b[5], uint64_t out[5]) { put_sc(out, sc_add(SC(a), SC(b))); }
void zo_sc_sub(const uint64_t a[5], const uint64_t b[5], uint64_t out[5]) { put_sc(out, sc_sub(SC(a), SC(b))); }
void zo_sc_neg(const uint64_t a[5], uint64_t out[5]) { put_sc(out, sc_neg(SC(a))); }
void zo_sc_mul(const uint64_t a[5], const uint64_t b[5], uint64_t out[5]) { put_sc(out, sc_mul(SC(a), SC(b))); }
void zo_sc_square(const uint64_t a[5], uint64_t out[5]) { put_sc(out, sc_square(SC(a))); }
void zo_sc_montgomery_mul(const uint64_t a[5], const uint64_t b[5], uint64_t out[5]) { put_sc(out, sc_montgomery_mul(SC(a), SC(b))); }
void zo_sc_to_montgomery(const uint64_t a[5], uint64_t out[5]) { put_sc(out, sc_to_montgomery(SC(a))); }
void zo_sc_from_montgomery(const uint64_t a[5], uint64_t out[5]) { put_sc(out, sc_from_montgomery(SC(a))); }
int zo_sc_from_bytes(const uint8_t bytes[32], uint64_t out[5]) { zo_sc r; int e = sc_from_bytes(bytes, &r); put_sc(out, r); return e; }
void zo_sc_to_bytes(const uint64_t a[5], uint8_t out[32]) { limbs_to_bytes(a, out); }
void zo_sc_half_without_mod(const uint64_t a[5], uint64_t out[5]) { put_sc(out, sc_half_without_mod(SC(a))); }
void zo_sc_half(const uint64_t a[5], uint64_t out[5]) { put_sc(out, sc_half(SC(a))); }
void zo_sc_shr(const uint64_t a[5], unsigned n, uint64_t out[5]) { put_sc(out, sc_shr(SC(a), n)); }
void zo_sc_pow(const uint64_t a[5], const uint64_t e[5], uint64_t out[5]) { put_sc(out, sc_pow(SC(a), SC(e))); }
void zo_sc_two_pow_k(uint64_t k, uint64_t out[5]) { limbs_two_pow(k, out); }
void zo_sc_from_i8(int8_t x, uint64_t out[5]) { put_sc(out, sc_from_i8(x)); }
void zo_sc_into_bits(const uint64_t a[5], uint8_t bits[256]) { sc_into_bits(SC(a), bits); }
void zo_sc_compute_naf(const uint64_t a[5], int8_t naf[256]) { sc_compute_naf(SC(a), naf); }
void zo_sc_compute_window_naf(const uint64_t a[5], unsigned width, int8_t naf[256]) { sc_compute_window_naf(SC(a), width, naf); }

void zo_pt_identity(uint64_t out[20]) { put_pt(out, PT_IDENTITY); }
void zo_pt_neg(const uint64_t p[20], uint64_t out[20]) { put_pt(out, pt_neg(PT(p))); }
void zo_pt_add(const uint64_t p[20], const uint64_t q[20], uint64_t out[20]) { put_pt(out, pt_add(PT(p), PT(q))); }
void zo_pt_sub(const uint64_t p[20], const uint64_t q[20], uint64_t out[20]) { put_pt(out, pt_sub(PT(p), PT(q))); }
void zo_pt_double(const uint64_t p[20], uint64_t out[20]) { put_pt(out, pt_double(PT(p))); }
void zo_pt_double_and_add(const uint64_t p[20], const uint64_t s[5], uint64_t out[20]) { put_pt(out, pt_double_and_add(PT(p), SC(s))); }
void zo_pt_ltr_bin_mul(const uint64_t p[20], const uint64_t s[5], uint64_t out[20]) { put_pt(out, pt_ltr_bin_mul(PT(p), SC(s))); }
void zo_pt_binary_naf_mul(const uint64_t p[20], const uint64_t s[5], uint64_t out[20]) { put_pt(out, pt_binary_naf_mul(PT(p), SC(s))); }
int zo_pt_to_affine(const uint64_t p[20], uint64_t xy[10]) {
    zo_fe x, y;
    if (pt_to_affine(PT(p), &x, &y)) return 1;
    put_fe(xy, x);
    put_fe(xy + 5, y);
    return 0;
}
int zo_pt_eq(const uint64_t p[20], const uint64_t q[20]) { return pt_eq(PT(p), PT(q)); }
int zo_pt_is_valid(const uint64_t p[20]) { return pt_is_valid(PT(p)); }
int zo_pt_new_from_y_coord(const uint64_t y[5], int sign, uint64_t out[20]) { zo_pt r; if (!pt_new_from_y_coord(FE(y), sign, &r)) return 0; put_pt(out, r); return 1; }
int zo_pt_compress(const uint64_t p[20], uint8_t out[32]) { return pt_compress(PT(p), out); }
int zo_pt_decompress(const uint8_t in[32], uint64_t out[20]) { zo_pt r; if (!pt_decompress(in, &r)) return 0; put_pt(out, r); return 1; }

int zo_ris_eq(const uint64_t p[20], const uint64_t q[20]) { return ris_eq(PT(p), PT(q)); }
void zo_ris_compress(const uint64_t p[20], uint8_t out[32]) { ris_compress(PT(p), out); }
int zo_ris_decompress(const uint8_t in[32], uint64_t out[20]) { zo_pt r; if (!ris_decompress(in, &r)) return 0; put_pt(out, r); return 1; }
void zo_ris_elligator(const uint64_t r0[5], uint64_t out[20]) { put_pt(out, ris_elligator(FE(r0))); }
void zo_ris_from_uniform_bytes(const uint8_t in[64], uint64_t out[20]) { put_pt(out, ris_from_uniform_bytes(in)); }

/* ------------------------------------------------------------------------------------------- */
/* Batch drivers                                                                                */
/* ------------------------------------------------------------------------------------------- */
typedef void (*range_fn)(void *ctx, size_t lo, size_t hi, int tid);
typedef struct { range_fn fn; void *ctx; size_t lo, hi; int tid; } job_t;
static void *job_main(void *arg) { job_t *j = (job_t *)arg; j->fn(j->ctx, j->lo, j->hi, j->tid); return NULL; }

static void parallel_for(size_t n, int threads, range_fn fn, void *ctx) {
    if (threads <= 1 || n < 2) { fn(ctx, 0, n, 0); return; }
    if ((size_t)threads > n) threads = (int)n;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)threads);
    size_t chunk = (n + (size_t)threads - 1) / (size_t)threads;
    for (int t = 0; t < threads; t++) {
        size_t lo = (size_t)t * chunk, hi = lo + chunk;
        if (lo > n) lo = n;
        if (hi > n) hi = n;
        jobs[t] = (job_t){fn, ctx, lo, hi, t};
        pthread_create(&th[t], NULL, job_main, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

typedef struct { const uint64_t *a, *b; uint64_t *o, *o2; const uint8_t *ib; uint8_t *ob; zo_pt *partials; } bctx;

#define BATCH2(name, stride_in, stride_out, expr)                                                      \
    static void name##_range(void *c, size_t lo, size_t hi, int tid) {                                  \
        (void)tid; bctx *x = (bctx *)c;                                                                 \
        for (size_t i = lo; i < hi; i++) {                                                              \
            const uint64_t *a = x->a + i * stride_in, *b = x->b ? x->b + i * stride_in : NULL;          \
            uint64_t *o = x->o + i * stride_out; (void)b;                                               \
            expr;                                                                                       \
        }                                                                                               \
    }

BATCH2(b_fe_mul, 5, 5, put_fe(o, fe_mul(FE(a), FE(b))))
BATCH2(b_fe_square, 5, 5, put_fe(o, fe_square(FE(a))))
BATCH2(b_fe_add, 5, 5, put_fe(o, fe_add(FE(a), FE(b))))
BATCH2(b_fe_sub, 5, 5, put_fe(o, fe_sub(FE(a), FE(b))))
BATCH2(b_fe_neg, 5, 5, put_fe(o, fe_neg(FE(a))))
BATCH2(b_fe_mul_square, 5, 5, (put_fe(o, fe_mul(FE(a), FE(b))), put_fe(x->o2 + i * 5, fe_square(FE(a)))))
BATCH2(b_sc_mul, 5, 5, put_sc(o, sc_mul(SC(a), SC(b))))
BATCH2(b_sc_square, 5, 5, put_sc(o, sc_square(SC(a))))
BATCH2(b_sc_add, 5, 5, put_sc(o, sc_add(SC(a), SC(b))))
BATCH2(b_sc_sub, 5, 5, put_sc(o, sc_sub(SC(a), SC(b))))
BATCH2(b_pt_add, 20, 20, put_pt(o, pt_add(PT(a), PT(b))))
BATCH2(b_pt_sub, 20, 20, put_pt(o, pt_sub(PT(a), PT(b))))
BATCH2(b_pt_double, 20, 20, put_pt(o, pt_double(PT(a))))
BATCH2(b_pt_neg, 20, 20, put_pt(o, pt_neg(PT(a))))

static void b_smul_range(void *c, size_t lo, size_t hi, int tid) {
    (void)tid; bctx *x = (bctx *)c;
    for (size_t i = lo; i < hi; i++) put_pt(x->o + i * 20, pt_double_and_add(PT(x->a + i * 20), SC(x->b + i * 5)));
}
static void b_affine_range(void *c, size_t lo, size_t hi, int tid) {
    (void)tid; bctx *x = (bctx *)c;
    for (size_t i = lo; i < hi; i++) {
        zo_fe ax = FE_ZERO, ay = FE_ZERO;
        pt_to_affine(PT(x->a + i * 20), &ax, &ay);
        put_fe(x->o + i * 10, ax);
        put_fe(x->o + i * 10 + 5, ay);
    }
}
static void b_riscomp_range(void *c, size_t lo, size_t hi, int tid) {
    (void)tid; bctx *x = (bctx *)c;
    for (size_t i = lo; i < hi; i++) ris_compress(PT(x->a + i * 20), x->ob + i * 32);
}
static void b_msm_range(void *c, size_t lo, size_t hi, int tid) {
    bctx *x = (bctx *)c;
    zo_pt acc = PT_IDENTITY;
    for (size_t i = lo; i < hi; i++) {
        zo_pt t = pt_double_and_add(PT(x->a + i * 20), SC(x->b + i * 5));
        acc = pt_add(&acc, &t);
    }
    x->partials[tid] = acc;
}

#define RUN(rangefn, A, B, O, O2) do { bctx c = {A, B, O, O2, NULL, NULL, NULL}; parallel_for(n, threads, rangefn, &c); } while (0)
void zo_fe_mul_batch(const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n, int threads) { RUN(b_fe_mul_range, a, b, out, NULL); }
void zo_fe_square_batch(const uint64_t *a, uint64_t *out, size_t n, int threads) { RUN(b_fe_square_range, a, NULL, out, NULL); }
void zo_fe_add_batch(const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n, int threads) { RUN(b_fe_add_range, a, b, out, NULL); }
void zo_fe_sub_batch(const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n, int threads) { RUN(b_fe_sub_range, a, b, out, NULL); }
void zo_fe_neg_batch(const uint64_t *a, uint64_t *out, size_t n, int threads) { RUN(b_fe_neg_range, a, NULL, out, NULL); }
void zo_fe_mul_square_batch(const uint64_t *a, const uint64_t *b, uint64_t *prod, uint64_t *sq, size_t n, int threads) { RUN(b_fe_mul_square_range, a, b, prod, sq); }
void zo_sc_mul_batch(const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n, int threads) { RUN(b_sc_mul_range, a, b, out, NULL); }
void zo_sc_square_batch(const uint64_t *a, uint64_t *out, size_t n, int threads) { RUN(b_sc_square_range, a, NULL, out, NULL); }
void zo_sc_add_batch(const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n, int threads) { RUN(b_sc_add_range, a, b, out, NULL); }
void zo_sc_sub_batch(const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n, int threads) { RUN(b_sc_sub_range, a, b, out, NULL); }
void zo_pt_add_batch(const uint64_t *p, const uint64_t *q, uint64_t *out, size_t n, int threads) { RUN(b_pt_add_range, p, q, out, NULL); }
void zo_pt_sub_batch(const uint64_t *p, const uint64_t *q, uint64_t *out, size_t n, int threads) { RUN(b_pt_sub_range, p, q, out, NULL); }
void zo_pt_double_batch(const uint64_t *p, uint64_t *out, size_t n, int threads) { RUN(b_pt_double_range, p, NULL, out, NULL); }
void zo_pt_neg_batch(const uint64_t *p, uint64_t *out, size_t n, int threads) { RUN(b_pt_neg_range, p, NULL, out, NULL); }
void zo_pt_scalar_mul_batch(const uint64_t *p, const uint64_t *s, uint64_t *out, size_t n, int threads) { RUN(b_smul_range, p, s, out, NULL); }
void zo_pt_to_affine_batch(const uint64_t *p, uint64_t *xy, size_t n, int threads) { RUN(b_affine_range, p, NULL, xy, NULL); }
void zo_ris_compress_batch(const uint64_t *p, uint8_t *out, size_t n, int threads) {
    bctx c = {p, NULL, NULL, NULL, NULL, out, NULL};
    parallel_for(n, threads, b_riscomp_range, &c);
}

void zo_msm_naive(const uint64_t *p, const uint64_t *s, size_t n, int threads, uint64_t out[20]) {
    if (threads < 1) threads = 1;
    if (n > 0 && (size_t)threads > n) threads = (int)n;
    zo_pt *partials = (zo_pt *)malloc(sizeof(zo_pt) * (size_t)threads);
    for (int t = 0; t < threads; t++) partials[t] = PT_IDENTITY;
    bctx c = {p, s, NULL, NULL, NULL, NULL, partials};
    parallel_for(n, threads, b_msm_range, &c);
    zo_pt acc = partials[0];
    for (int t = 1; t < threads; t++) acc = pt_add(&acc, &partials[t]);
    put_pt(out, acc);
    free(partials);
}

/* ------------------------------------------------------------------------------------------- */
/* Synthetic inputs: SplitMix64 as a counter PRNG (word j of element i = mix(key + ctr*phi)).    */
/* Same construction as dusk_zerocaf_b200/synth.py; tests check the two agree.                   */
/* ------------------------------------------------------------------------------------------- */
static inline u64 splitmix(u64 seed, u64 stream, u64 ctr) {
    u64 z = seed + stream * 0xD1B54A32D192ED03ULL + (ctr + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static void synth_bytes(u64 seed, u64 stream, size_t i, uint8_t by[32]) {
    for (int j = 0; j < 4; j++) {
        u64 w = splitmix(seed, stream, 4 * (u64)i + (u64)j);
        for (int k = 0; k < 8; k++) by[8 * j + k] = (uint8_t)(w >> (8 * k));
    }
}
void zo_synth_fe(uint64_t seed, uint64_t stream, size_t first, size_t n, uint64_t *out) {
    for (size_t i = 0; i < n; i++) {
        uint8_t by[32];
        synth_bytes(seed, stream, first + i, by);
        by[31] &= 0x07; /* src/field.rs:138 */
        put_fe(out + 5 * i, fe_from_bytes(by));
    }
}
void zo_synth_scalar(uint64_t seed, uint64_t stream, size_t first, size_t n, uint64_t *out) {
    for (size_t i = 0; i < n; i++) {
        uint8_t by[32];
        zo_sc s;
        synth_bytes(seed, stream, first + i, by);
        by[31] &= 0x01; /* src/scalar.rs:107 */
        sc_from_bytes(by, &s);
        put_sc(out + 5 * i, s);
    }
}
