// zc_point.cuh -- extended twisted-Edwards points on the Sonny curve  -x^2 + y^2 = 1 + d x^2 y^2  (a = -1),
// d = -126296/126297 (reference constants.rs:86-92), one point per thread.
//
// Device-side replacement for /root/reference/src/edwards.rs:
//   EdwardsPoint {X,Y,Z,T} :336-342, Identity :381-391, Neg :440-455, Add :465-489, Sub :503-531,
//   Double (= self + self) :579-592, double_and_add :102-120.
// RistrettoPoint (ristretto.rs:157-158) is a newtype over EdwardsPoint whose Add/Sub/Double/Mul forward to these
// (ristretto.rs:248-392), so the same device functions serve both.
//
// "ref" functions evaluate exactly the reference's polynomials, so the (X:Y:Z:T) they return is the same
// projective representative, limb for limb.  "fast" functions (dedicated doubling, cached-operand addition) return
// the same group element with a different representative; callers compare those canonically.
#pragma once
#include "zc_fe.cuh"

namespace zc {

struct Pt { Fe X, Y, Z, T; };   // coordinates in Montgomery form unless stated otherwise

// d*R mod p, 2d*R mod p
__device__ __forceinline__ Fe D_MONT()  { return Fe{{0xa911bcacu, 0x4dc31488u, 0x96c021a0u, 0x150160ffu, 0xf2abb033u, 0x6960412fu, 0x953fedb5u, 0x0bcdf760u}}; }
__device__ __forceinline__ Fe D2_MONT() { return Fe{{0xf52da56bu, 0x4373c5f6u, 0x8a88a66au, 0x1523c820u, 0xe5576066u, 0xd2c0825fu, 0x2a7fdb6au, 0x079beec1u}}; }

__device__ __forceinline__ Pt pt_identity_mont() {   // (0, 1, 1, 0)  edwards.rs:381-391
  Fe z{{0, 0, 0, 0, 0, 0, 0, 0}};
  Fe one = Consts<ModP>::R1();
  return Pt{z, one, one, z};
}

__device__ __forceinline__ Pt pt_neg(const Pt& p) {   // (-X, Y, Z, -T)  edwards.rs:440-455
  return Pt{fe_neg<ModP>(p.X), p.Y, p.Z, fe_neg<ModP>(p.T)};
}

// ---- P + Q with the reference's formulas (edwards.rs:473-487), Montgomery-form in and out --------------
__device__ __forceinline__ Pt pt_add_ref(const Pt& p, const Pt& q) {
  typedef ModP M;
  Fe A = mont_mul<M>(p.X, q.X);
  Fe B = mont_mul<M>(p.Y, q.Y);
  Fe C = mont_mul<M>(mont_mul<M>(p.T, q.T), D_MONT());
  Fe D = mont_mul<M>(p.Z, q.Z);
  Fe E = mont_mul<M>(fe_add<M>(p.X, p.Y), fe_add<M>(q.X, q.Y));
  E = fe_sub<M>(fe_sub<M>(E, A), B);
  Fe F = fe_sub<M>(D, C);
  Fe G = fe_add<M>(D, C);
  Fe H = fe_add<M>(B, A);
  Pt r;
  r.X = mont_mul<M>(E, F);
  r.Y = mont_mul<M>(G, H);
  r.Z = mont_mul<M>(F, G);
  r.T = mont_mul<M>(E, H);
  return r;
}

// ---- P + Q, NORMAL-form in and out, still the reference's polynomials, 12 Montgomery products ----------
// With normal-form inputs every first-layer product carries a factor 1/R: A' = A/R ... H' = H/R.  Lifting only
// F' and H' by R^3 (one product with R^4 each) makes the four output products come out exact:
//   mont(E/R, F R^2) = EF,  mont(G/R, H R^2) = GH,  mont(F R^2, G/R) = FG,  mont(E/R, H R^2) = EH.
__device__ __forceinline__ Pt pt_add_ref_normal(const Pt& p, const Pt& q) {
  typedef ModP M;
  Fe A = mont_mul<M>(p.X, q.X);
  Fe B = mont_mul<M>(p.Y, q.Y);
  Fe C = mont_mul<M>(mont_mul<M>(p.T, q.T), D_MONT());
  Fe D = mont_mul<M>(p.Z, q.Z);
  Fe E = mont_mul<M>(fe_add<M>(p.X, p.Y), fe_add<M>(q.X, q.Y));
  E = fe_sub<M>(fe_sub<M>(E, A), B);
  Fe F = fe_sub<M>(D, C);
  Fe G = fe_add<M>(D, C);
  Fe H = fe_add<M>(B, A);
  const Fe R4 = Consts<M>::R4();
  F = mont_mul<M>(F, R4);
  H = mont_mul<M>(H, R4);
  Pt r;
  r.X = mont_mul<M>(E, F);
  r.Y = mont_mul<M>(G, H);
  r.Z = mont_mul<M>(F, G);
  r.T = mont_mul<M>(E, H);
  return r;
}

// ---- dedicated doubling, dbl-2008-hwcd with a = -1 (4M + 4S): same group element as P + P, other representative
__device__ __forceinline__ Pt pt_double_fast(const Pt& p) {
  typedef ModP M;
  Fe A = mont_sqr<M>(p.X);
  Fe B = mont_sqr<M>(p.Y);
  Fe Z2 = mont_sqr<M>(p.Z);
  Fe C = fe_add<M>(Z2, Z2);
  Fe D = fe_neg<M>(A);                       // a*A, a = -1
  Fe S = fe_add<M>(p.X, p.Y);
  Fe E = fe_sub<M>(fe_sub<M>(mont_sqr<M>(S), A), B);
  Fe G = fe_add<M>(D, B);
  Fe F = fe_sub<M>(G, C);
  Fe H = fe_sub<M>(D, B);
  Pt r;
  r.X = mont_mul<M>(E, F);
  r.Y = mont_mul<M>(G, H);
  r.Z = mont_mul<M>(F, G);
  r.T = mont_mul<M>(E, H);
  return r;
}

// ---- cached operand for repeated additions: (Y+X, Y-X, Z, 2dT)  -> add costs 8M (7M when Z == 1) --------
struct PtCached { Fe YpX, YmX, Z, T2d; };

__device__ __forceinline__ PtCached pt_to_cached(const Pt& p) {
  typedef ModP M;
  return PtCached{fe_add<M>(p.Y, p.X), fe_sub<M>(p.Y, p.X), p.Z, mont_mul<M>(p.T, D2_MONT())};
}
__device__ __forceinline__ PtCached pt_cached_neg(const PtCached& c) {
  return PtCached{c.YmX, c.YpX, c.Z, fe_neg<ModP>(c.T2d)};
}

// add-2008-hwcd-3 (a = -1):  A=(Y1-X1)(Y2-X2) B=(Y1+X1)(Y2+X2) C=T1*2d*T2 D=2*Z1*Z2
__device__ __forceinline__ Pt pt_add_cached(const Pt& p, const PtCached& q) {
  typedef ModP M;
  Fe A = mont_mul<M>(fe_sub<M>(p.Y, p.X), q.YmX);
  Fe B = mont_mul<M>(fe_add<M>(p.Y, p.X), q.YpX);
  Fe C = mont_mul<M>(p.T, q.T2d);
  Fe D = mont_mul<M>(p.Z, q.Z);
  D = fe_add<M>(D, D);
  Fe E = fe_sub<M>(B, A);
  Fe F = fe_sub<M>(D, C);
  Fe G = fe_add<M>(D, C);
  Fe H = fe_add<M>(B, A);
  Pt r;
  r.X = mont_mul<M>(E, F);
  r.Y = mont_mul<M>(G, H);
  r.Z = mont_mul<M>(F, G);
  r.T = mont_mul<M>(E, H);
  return r;
}

// cached operand with Z == 1 (affine): (y+x, y-x, 2d*x*y), 7M per add
struct PtAffCached { Fe YpX, YmX, T2d; };

__device__ __forceinline__ Pt pt_add_affcached(const Pt& p, const PtAffCached& q, bool negate) {
  typedef ModP M;
  Fe qa = negate ? q.YpX : q.YmX;
  Fe qb = negate ? q.YmX : q.YpX;
  Fe A = mont_mul<M>(fe_sub<M>(p.Y, p.X), qa);
  Fe B = mont_mul<M>(fe_add<M>(p.Y, p.X), qb);
  Fe C = mont_mul<M>(p.T, q.T2d);
  Fe D = fe_add<M>(p.Z, p.Z);
  Fe E = fe_sub<M>(B, A);
  Fe F = negate ? fe_add<M>(D, C) : fe_sub<M>(D, C);
  Fe G = negate ? fe_sub<M>(D, C) : fe_add<M>(D, C);
  Fe H = fe_add<M>(B, A);
  Pt r;
  r.X = mont_mul<M>(E, F);
  r.Y = mont_mul<M>(G, H);
  r.Z = mont_mul<M>(F, G);
  r.T = mont_mul<M>(E, H);
  return r;
}

// ---- generic add of two extended points, fast formulas (9M): same element as pt_add_ref, other representative
__device__ __forceinline__ Pt pt_add_fast(const Pt& p, const Pt& q) {
  return pt_add_cached(p, pt_to_cached(q));
}

// ---- load / store at the ABI layout: [u64;20] = X|Y|Z|T, radix 2^52 (edwards.rs:336-342) ---------------
__device__ __forceinline__ Pt pt_load52(const uint64_t* __restrict__ p) {
  return Pt{fe_load52(p), fe_load52(p + 5), fe_load52(p + 10), fe_load52(p + 15)};
}
__device__ __forceinline__ void pt_store52(uint64_t* __restrict__ p, const Pt& a) {
  fe_store52(p, a.X); fe_store52(p + 5, a.Y); fe_store52(p + 10, a.Z); fe_store52(p + 15, a.T);
}
__device__ __forceinline__ Pt pt_to_mont(const Pt& a) {
  return Pt{to_mont<ModP>(a.X), to_mont<ModP>(a.Y), to_mont<ModP>(a.Z), to_mont<ModP>(a.T)};
}
__device__ __forceinline__ Pt pt_from_mont(const Pt& a) {
  return Pt{from_mont<ModP>(a.X), from_mont<ModP>(a.Y), from_mont<ModP>(a.Z), from_mont<ModP>(a.T)};
}

}  // namespace zc
