// zc_quad.cuh -- lazy linear combinations and the four-lanes-per-point operations of the MSM tail (window chain,
// reduction trees): a point lives in four consecutive lanes, lane q holds coordinate q of (X, Y, Z, T).
// Shared by zc_msm.cu and tools/ubench/chainbench.cu.
#pragma once
#include "zc_point.cuh"

namespace zc {

// 32-byte field elements / 128-byte points as 8 / 32 consecutive u32 words (16-byte aligned)
__device__ __forceinline__ void ld_fe(const uint32_t* __restrict__ p, Fe& a) {
  uint4 lo = *reinterpret_cast<const uint4*>(p);
  uint4 hi = *reinterpret_cast<const uint4*>(p + 4);
  a.w[0] = lo.x; a.w[1] = lo.y; a.w[2] = lo.z; a.w[3] = lo.w;
  a.w[4] = hi.x; a.w[5] = hi.y; a.w[6] = hi.z; a.w[7] = hi.w;
}
__device__ __forceinline__ void st_fe(uint32_t* __restrict__ p, const Fe& a) {
  *reinterpret_cast<uint4*>(p) = make_uint4(a.w[0], a.w[1], a.w[2], a.w[3]);
  *reinterpret_cast<uint4*>(p + 4) = make_uint4(a.w[4], a.w[5], a.w[6], a.w[7]);
}
__device__ __forceinline__ Pt ld_pt(const uint32_t* __restrict__ p) {
  Pt r; ld_fe(p, r.X); ld_fe(p + 8, r.Y); ld_fe(p + 16, r.Z); ld_fe(p + 24, r.T); return r;
}
__device__ __forceinline__ void st_pt(uint32_t* __restrict__ p, const Pt& a) {
  st_fe(p, a.X); st_fe(p + 8, a.Y); st_fe(p + 16, a.Z); st_fe(p + 24, a.T);
}

// lazy linear combinations (bucket accumulation, window chain, reduction trees): no conditional subtraction, results < 4m (inputs canonical);
// mont_mul accepts them because the product of any two stays below R m = 2^256 m (16 m^2 > 8 m^2).
__device__ __forceinline__ Fe fe_dbl_lazy(const Fe& a) {                 // 2a < 2m
  Fe r;
#pragma unroll
  for (int k = 7; k > 0; k--) r.w[k] = __funnelshift_l(a.w[k - 1], a.w[k], 1);
  r.w[0] = a.w[0] << 1;
  return r;
}
// a - b + K m  (K <= 8), b < K m
template <int K>
__device__ __forceinline__ Fe fe_sub_lazy(const Fe& a, const Fe& b) {
  typedef ModP M;
  constexpr uint64_t m01 = ((uint64_t)M::M1 << 32 | M::M0), m23 = ((uint64_t)M::M3 << 32 | M::M2);
  // K * m as words: the four low words times K spill into word 4 (k4)
  constexpr unsigned __int128 lowK = ((unsigned __int128)m23 << 64 | m01) * K;
  constexpr uint32_t k0 = (uint32_t)lowK, k1 = (uint32_t)(lowK >> 32), k2 = (uint32_t)(lowK >> 64), k3 = (uint32_t)(lowK >> 96),
                     k4 = 0u, k7 = M::M7 * K;      // c < 2^125: K c < 2^128 for K <= 8
  static_assert(K >= 1 && K <= 8, "K m must fit 256 bits");
  Fe r;
  asm("add.cc.u32  %0, %8,  %16;\n\t"
      "addc.cc.u32 %1, %9,  %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, 0;\n\t"
      "addc.cc.u32 %6, %14, 0;\n\t"
      "addc.u32    %7, %15, %21;\n\t"
      "sub.cc.u32  %0, %0, %22;\n\t"
      "subc.cc.u32 %1, %1, %23;\n\t"
      "subc.cc.u32 %2, %2, %24;\n\t"
      "subc.cc.u32 %3, %3, %25;\n\t"
      "subc.cc.u32 %4, %4, %26;\n\t"
      "subc.cc.u32 %5, %5, %27;\n\t"
      "subc.cc.u32 %6, %6, %28;\n\t"
      "subc.u32    %7, %7, %29;\n\t"
      : "=&r"(r.w[0]), "=&r"(r.w[1]), "=&r"(r.w[2]), "=&r"(r.w[3]), "=&r"(r.w[4]), "=&r"(r.w[5]), "=&r"(r.w[6]), "=&r"(r.w[7])
      : "r"(a.w[0]), "r"(a.w[1]), "r"(a.w[2]), "r"(a.w[3]), "r"(a.w[4]), "r"(a.w[5]), "r"(a.w[6]), "r"(a.w[7]),
        "r"(k0), "r"(k1), "r"(k2), "r"(k3), "r"(k4), "r"(k7),
        "r"(b.w[0]), "r"(b.w[1]), "r"(b.w[2]), "r"(b.w[3]), "r"(b.w[4]), "r"(b.w[5]), "r"(b.w[6]), "r"(b.w[7]));
  return r;
}
// a + b without the conditional subtraction (a, b < m: the sum is < 2m < 2^254)
__device__ __forceinline__ Fe fe_add_lazy(const Fe& a, const Fe& b) {
  Fe r;
  asm("add.cc.u32  %0, %8,  %16;\n\t"
      "addc.cc.u32 %1, %9,  %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32    %7, %15, %23;\n\t"
      : "=&r"(r.w[0]), "=&r"(r.w[1]), "=&r"(r.w[2]), "=&r"(r.w[3]), "=&r"(r.w[4]), "=&r"(r.w[5]), "=&r"(r.w[6]), "=&r"(r.w[7])
      : "r"(a.w[0]), "r"(a.w[1]), "r"(a.w[2]), "r"(a.w[3]), "r"(a.w[4]), "r"(a.w[5]), "r"(a.w[6]), "r"(a.w[7]),
        "r"(b.w[0]), "r"(b.w[1]), "r"(b.w[2]), "r"(b.w[3]), "r"(b.w[4]), "r"(b.w[5]), "r"(b.w[6]), "r"(b.w[7]));
  return r;
}
// lane-to-lane copies of a field element / point
__device__ __forceinline__ Fe shfl_fe(const Fe& a, int src) {
  Fe r;
#pragma unroll
  for (int k = 0; k < 8; k++) r.w[k] = __shfl_sync(0xffffffffu, a.w[k], src);
  return r;
}
// ---- four lanes per point operation (window chain and the latency-bound reduction kernels) ---------------------------
// Lane q = lane & 3 holds coordinate q (X, Y, Z, T) of the running point; the four independent field multiplications
// of each of the two stages of a doubling / addition run on the four lanes, the operands travel by shuffle.
//
// The serial tail of a sharded MSM is a few hundred dependent four-lane operations, so what counts is the number of
// dependent field multiplications and the instructions between them (tools/ubench/chainbench.cu measures both):
//   * per-lane operand selection is branch-free (selp): written as nested ?: on the lane index it compiled to divergent
//     branches, eight reconvergence regions per operation, 1300 of the 2900 cycles of a doubling;
//   * nothing is canonicalised: coordinates stay lazily reduced (< 2.5 m) between operations, every linear combination
//     is a plain carry chain and mont_mul_lazy skips the conditional subtraction;
//   * 2d is taken out of the addition: d = -126296/126297 (constants.rs:86-92), so with A, B, D scaled by 126297 and
//     C = -252592 T1 T2 the four outputs are the projective point scaled by 126297^2 -- two multiplications by 18-bit
//     constants (one row of wide multiplies + a fold with 2^252 = -c) replace the full 2d T2 product, and an addition
//     is two dependent multiplications deep instead of three.
__device__ __forceinline__ uint32_t selp_u32(uint32_t a, uint32_t b, int p) {     // p ? a : b, never a branch
  uint32_t r;
  asm("{\n\t.reg .pred sp;\n\tsetp.ne.s32 sp, %3, 0;\n\tselp.b32 %0, %1, %2, sp;\n\t}" : "=r"(r) : "r"(a), "r"(b), "r"(p));
  return r;
}
// the stage-2 operands of lane q:  X3 = E F, Y3 = G H, Z3 = F G, T3 = E H
__device__ __forceinline__ void quad_pick(Fe& u, Fe& v, const Fe& E, const Fe& F, const Fe& G, const Fe& H, int q) {
  const int q0 = q == 0, q1 = q == 1, q2 = q == 2;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    u.w[k] = selp_u32(G.w[k], selp_u32(F.w[k], E.w[k], q2), q1);       // E G F E
    v.w[k] = selp_u32(F.w[k], selp_u32(G.w[k], H.w[k], q2), q0);       // F H G H
  }
}
// coordinate q of a point every lane holds in full
__device__ __forceinline__ Fe pt_coord(const Pt& p, int q) {
  Fe r;
  const int q0 = q == 0, q1 = q == 1, q2 = q == 2;
#pragma unroll
  for (int k = 0; k < 8; k++) r.w[k] = selp_u32(p.X.w[k], selp_u32(p.Y.w[k], selp_u32(p.Z.w[k], p.T.w[k], q2), q1), q0);
  return r;
}
// x * k mod p for x < 2^256, k < 2^18: result < 2m
__device__ __forceinline__ Fe fe_mul_small(const Fe& x, uint32_t k) {
  typedef ModP M;
  uint32_t ev[8], od[8], t[9];
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    asm("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=r"(ev[j]), "=r"(ev[j + 1]) : "r"(x.w[j]), "r"(k));
    asm("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=r"(od[j]), "=r"(od[j + 1]) : "r"(x.w[j + 1]), "r"(k));
  }
  t[0] = ev[0];
  asm("add.cc.u32  %0, %8,  %15;\n\t"
      "addc.cc.u32 %1, %9,  %16;\n\t"
      "addc.cc.u32 %2, %10, %17;\n\t"
      "addc.cc.u32 %3, %11, %18;\n\t"
      "addc.cc.u32 %4, %12, %19;\n\t"
      "addc.cc.u32 %5, %13, %20;\n\t"
      "addc.cc.u32 %6, %14, %21;\n\t"
      "addc.u32    %7, %22, 0;\n\t"
      : "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]), "=&r"(t[8])
      : "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]),
        "r"(od[0]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]));
  // t = lo + 2^252 hi,  hi < 2^22;  2^252 = -c (mod p):  x k = lo + m - hi c
  const uint32_t hi = (t[7] >> M::TOP) | (t[8] << (32 - M::TOP));
  t[7] &= (1u << M::TOP) - 1u;
  uint32_t e0, e1, e2, e3, o0, o1, o2, o3;
  asm("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=r"(e0), "=r"(e1) : "r"(hi), "r"(M::M0));
  asm("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=r"(o0), "=r"(o1) : "r"(hi), "r"(M::M1));
  asm("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=r"(e2), "=r"(e3) : "r"(hi), "r"(M::M2));
  asm("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=r"(o2), "=r"(o3) : "r"(hi), "r"(M::M3));
  uint32_t h1, h2, h3, h4;
  asm("add.cc.u32  %0, %4, %7;\n\t"
      "addc.cc.u32 %1, %5, %8;\n\t"
      "addc.cc.u32 %2, %6, %9;\n\t"
      "addc.u32    %3, %10, 0;\n\t"
      : "=&r"(h1), "=&r"(h2), "=&r"(h3), "=&r"(h4)
      : "r"(e1), "r"(e2), "r"(e3), "r"(o0), "r"(o1), "r"(o2), "r"(o3));
  Fe r;
  asm("add.cc.u32  %0, %8,  %16;\n\t"
      "addc.cc.u32 %1, %9,  %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, 0;\n\t"
      "addc.cc.u32 %5, %13, 0;\n\t"
      "addc.cc.u32 %6, %14, 0;\n\t"
      "addc.u32    %7, %15, %20;\n\t"
      "sub.cc.u32  %0, %0, %21;\n\t"
      "subc.cc.u32 %1, %1, %22;\n\t"
      "subc.cc.u32 %2, %2, %23;\n\t"
      "subc.cc.u32 %3, %3, %24;\n\t"
      "subc.cc.u32 %4, %4, %25;\n\t"
      "subc.cc.u32 %5, %5, 0;\n\t"
      "subc.cc.u32 %6, %6, 0;\n\t"
      "subc.u32    %7, %7, 0;\n\t"
      : "=&r"(r.w[0]), "=&r"(r.w[1]), "=&r"(r.w[2]), "=&r"(r.w[3]), "=&r"(r.w[4]), "=&r"(r.w[5]), "=&r"(r.w[6]), "=&r"(r.w[7])
      : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]),
        "r"(M::M0), "r"(M::M1), "r"(M::M2), "r"(M::M3), "r"(M::M7),
        "r"(e0), "r"(h1), "r"(h2), "r"(h3), "r"(h4));
  return r;
}
constexpr uint32_t EDW_D_DEN = 126297u, EDW_2D_NUM = 252592u;      // 2d = -EDW_2D_NUM / EDW_D_DEN

// X3 = E F, Y3 = G H, Z3 = F G, T3 = E H without the conditional subtraction
__device__ __forceinline__ Fe quad_stage2_lazy(const Fe& E, const Fe& F, const Fe& G, const Fe& H, int q) {
  Fe u, v;
  quad_pick(u, v, E, F, G, H, q);
  return mont_mul_lazy<ModP>(u, v);
}
// coordinates < 2.5 m in, < 2.5 m out (stage 1 < 1.4 m; E, C < 2.8 m, G < 3.4 m, F < 6.4 m, H <= 3 m; products < 21.7 m^2)
__device__ __forceinline__ Fe quad_double_inl(const Fe& c, int q, int qbase) {
  typedef ModP M;
  const Fe z = shfl_fe(c, qbase + 2);
  Fe in2;
  const int q3 = q == 3;
#pragma unroll
  for (int k = 0; k < 8; k++) in2.w[k] = selp_u32(z.w[k], c.w[k], q3);     // X^2, Y^2, Z^2 and T Z (= X Y: E = 2 T Z)
  const Fe s = mont_mul_lazy<M>(c, in2);
  const Fe A = shfl_fe(s, qbase), B = shfl_fe(s, qbase + 1), ZZ = shfl_fe(s, qbase + 2), TZ = shfl_fe(s, qbase + 3);
  const Fe E = fe_dbl_lazy(TZ);
  const Fe C = fe_dbl_lazy(ZZ);
  const Fe G = fe_sub_lazy<2>(B, A);
  const Fe F = fe_sub_lazy<3>(G, C);
  const Fe zero{{0, 0, 0, 0, 0, 0, 0, 0}};
  const Fe H = fe_sub_lazy<3>(zero, fe_add_lazy(A, B));
  return quad_stage2_lazy(E, F, G, H, q);
}
// c (distributed, < 2.5 m) += p (every lane holds all of p, coordinates < 2.5 m); result < 2 m
__device__ __forceinline__ Fe quad_add_inl(const Fe& c, const Pt& p, int q, int qbase) {
  typedef ModP M;
  const Fe x1 = shfl_fe(c, qbase), y1 = shfl_fe(c, qbase + 1);
  const Fe d1 = fe_sub_lazy<3>(y1, x1), s1 = fe_add_lazy(y1, x1);          // < 5.5 m, < 5 m
  const Fe d2 = fe_sub_lazy<3>(p.Y, p.X), s2 = fe_add_lazy(p.Y, p.X);
  const Fe z2 = fe_dbl_lazy(p.Z);
  Fe u, v;
  const int q0 = q == 0, q1 = q == 1, q2 = q == 2;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    u.w[k] = selp_u32(d1.w[k], selp_u32(s1.w[k], c.w[k], q1), q0);
    v.w[k] = selp_u32(d2.w[k], selp_u32(s2.w[k], selp_u32(z2.w[k], p.T.w[k], q2), q1), q0);
  }
  Fe s = mont_mul_lazy<M>(u, v);                                            // A, B, 2 Z1 Z2, T1 T2: < 2.9 m
  s = fe_mul_small(s, selp_u32(EDW_2D_NUM, EDW_D_DEN, q == 3));                     // 126297 (A, B, D), 252592 T1 T2 = -126297 C: < 2 m
  const Fe A = shfl_fe(s, qbase), B = shfl_fe(s, qbase + 1), D = shfl_fe(s, qbase + 2), Cn = shfl_fe(s, qbase + 3);
  const Fe E = fe_sub_lazy<2>(B, A);
  const Fe F = fe_add_lazy(D, Cn);                                          // D - C
  const Fe G = fe_sub_lazy<2>(D, Cn);                                       // D + C
  const Fe H = fe_add_lazy(B, A);
  return quad_stage2_lazy(E, F, G, H, q);                                   // < 16 m^2 / R + m = 2 m
}

// out-of-line copies for kernels with several call sites (one body in the instruction cache)
static __device__ __noinline__ Fe quad_double(Fe c, int q, int qbase) { return quad_double_inl(c, q, qbase); }
static __device__ __noinline__ Fe quad_add(Fe c, Pt p, int q, int qbase) { return quad_add_inl(c, p, q, qbase); }

// x < 4m -> canonical
__device__ __forceinline__ Fe fe_canon4(Fe x) {
  reduce_once<ModP>(x); reduce_once<ModP>(x); reduce_once<ModP>(x);
  return x;
}

}  // namespace zc
