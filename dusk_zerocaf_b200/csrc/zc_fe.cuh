// zc_fe.cuh -- 253-bit modular arithmetic for sm_100a, one residue per thread.
//
// Device-side replacement for the reference's u64 backend
//   /root/reference/src/backend/u64/field.rs   (FieldElement: Add :191-207, Sub :217-240, Mul :250-262,
//                                               Square :302-315, montgomery_reduce :780-813)
//   /root/reference/src/backend/u64/scalar.rs  (Scalar:       Add :184-200, Sub :210-237, Mul :247-258,
//                                               Square :272-283, montgomery_reduce :617-652)
//
// The reference keeps values in NORMAL form as 5 x 52-bit limbs and Montgomery-reduces every product twice
// (R = 2^260).  Here a residue lives in registers as 8 x u32 words (= 4 x u64) in MONTGOMERY form with
// R = 2^256, so a chain of multiplications pays one reduction each.  Only canonical values are observable
// at the ABI, so the results are bit-identical to the reference's.
//
// Both moduli have the shape  m = 2^k + c  with c < 2^125:  words 4,5,6 of m are zero and word 7 is a single
// bit, so the "m * modulus" half of the Montgomery step costs 4 wide multiplies plus two shifts instead of 8.
//
// The multiply is a CIOS loop, fully unrolled, written as PTX mad.lo.cc / madc.hi.cc carry chains: even
// and odd words of `a` feed two independent chains per row so the 64-bit products never overlap.
#pragma once
#include <stdint.h>

namespace zc {

// p = 2^252 + 27742317777372353535851937790883648493   (constants.rs:30-36 FIELD_L)
struct ModP {
  static constexpr uint32_t M0 = 0x5cf5d3edu, M1 = 0x5812631au, M2 = 0xa2f79cd6u, M3 = 0x14def9deu;
  static constexpr uint32_t M7 = 0x10000000u;   // 2^252 = 2^(224+28)
  static constexpr int TOP = 28;
  static constexpr uint32_t NINV = 0x12547e1bu;  // -p^-1 mod 2^32
};
// L = 2^249 + 14490550575682688738086195780655237219    (constants.rs:9 L)
struct ModL {
  static constexpr uint32_t M0 = 0x755fc863u, M1 = 0x6ab4036fu, M2 = 0x822fd593u, M3 = 0x0ae6c74du;
  static constexpr uint32_t M7 = 0x02000000u;   // 2^249 = 2^(224+25)
  static constexpr int TOP = 25;
  static constexpr uint32_t NINV = 0x84a706b5u;
};

struct Fe { uint32_t w[8]; };   // little-endian 32-bit words; Montgomery or normal form by context

template <class M> struct Consts;
template <> struct Consts<ModP> {
  // R mod p (Montgomery one), R^2 mod p, R^3, R^4
  static __device__ __forceinline__ Fe R1() { return Fe{{0x8d98951du, 0xd6ec3174u, 0x737dcf70u, 0xc6ef5bf4u, 0xfffffffeu, 0xffffffffu, 0xffffffffu, 0x0fffffffu}}; }
  static __device__ __forceinline__ Fe R2() { return Fe{{0x449c0f01u, 0xa40611e3u, 0x68859347u, 0xd00e1ba7u, 0x17f5be65u, 0xceec73d2u, 0x7c309a3du, 0x0399411bu}}; }
  static __device__ __forceinline__ Fe R3() { return Fe{{0x7b83a2dbu, 0x2a9e4968u, 0xaef7f3ecu, 0x278324e6u, 0x04ec5b65u, 0x8065dc6cu, 0x3599cec7u, 0x0e530b77u}}; }
  static __device__ __forceinline__ Fe R4() { return Fe{{0x42419a0du, 0x3a3dc222u, 0x023493f7u, 0x9d31cab2u, 0xb6870058u, 0xe7faf80eu, 0x5a45ffd7u, 0x09dc924eu}}; }
};
template <> struct Consts<ModL> {
  static __device__ __forceinline__ Fe R1() { return Fe{{0xc57b96e3u, 0x10b24bb4u, 0x6a450bdeu, 0x9783208cu, 0xfffffffau, 0xffffffffu, 0xffffffffu, 0x01ffffffu}}; }
  static __device__ __forceinline__ Fe R2() { return Fe{{0x050c31b2u, 0xfcbbafc0u, 0x48fd51d3u, 0x0d536753u, 0x98d542e5u, 0x0509b170u, 0xd0a04e90u, 0x01e73226u}}; }
  static __device__ __forceinline__ Fe R3() { return Fe{{0xbe881578u, 0x310d3917u, 0xa84b6a74u, 0xbac1268bu, 0xe97c1665u, 0x5efa1d46u, 0x56b447a0u, 0x01053032u}}; }
  static __device__ __forceinline__ Fe R4() { return Fe{{0x364b0effu, 0xfeedb71eu, 0x5c9e4192u, 0x28af3a50u, 0x79210933u, 0x5a1d5183u, 0xdac3a770u, 0x017c5642u}}; }
};

// ---- conditional subtraction of the modulus: r = (t >= m) ? t - m : t ---------------------------------
template <class M>
__device__ __forceinline__ void reduce_once(Fe& t) {
  uint32_t s0, s1, s2, s3, s4, s5, s6, s7, bw;
  asm("sub.cc.u32  %0, %9,  %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, 0;\n\t"
      "subc.cc.u32 %5, %14, 0;\n\t"
      "subc.cc.u32 %6, %15, 0;\n\t"
      "subc.cc.u32 %7, %16, %21;\n\t"
      "subc.u32    %8, 0, 0;\n\t"
      : "=r"(s0), "=r"(s1), "=r"(s2), "=r"(s3), "=r"(s4), "=r"(s5), "=r"(s6), "=r"(s7), "=r"(bw)
      : "r"(t.w[0]), "r"(t.w[1]), "r"(t.w[2]), "r"(t.w[3]), "r"(t.w[4]), "r"(t.w[5]), "r"(t.w[6]), "r"(t.w[7]),
        "r"(M::M0), "r"(M::M1), "r"(M::M2), "r"(M::M3), "r"(M::M7));
  const bool keep = (bw != 0);   // borrow -> t < m -> keep t
  t.w[0] = keep ? t.w[0] : s0; t.w[1] = keep ? t.w[1] : s1; t.w[2] = keep ? t.w[2] : s2; t.w[3] = keep ? t.w[3] : s3;
  t.w[4] = keep ? t.w[4] : s4; t.w[5] = keep ? t.w[5] : s5; t.w[6] = keep ? t.w[6] : s6; t.w[7] = keep ? t.w[7] : s7;
}

// ---- r = a + b mod m  (a, b canonical)          field.rs:191-207 / scalar.rs:184-200 -------------------
template <class M>
__device__ __forceinline__ Fe fe_add(const Fe& a, const Fe& b) {
  Fe r;
  asm("add.cc.u32  %0, %8,  %16;\n\t"
      "addc.cc.u32 %1, %9,  %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32    %7, %15, %23;\n\t"
      : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
      : "r"(a.w[0]), "r"(a.w[1]), "r"(a.w[2]), "r"(a.w[3]), "r"(a.w[4]), "r"(a.w[5]), "r"(a.w[6]), "r"(a.w[7]),
        "r"(b.w[0]), "r"(b.w[1]), "r"(b.w[2]), "r"(b.w[3]), "r"(b.w[4]), "r"(b.w[5]), "r"(b.w[6]), "r"(b.w[7]));
  reduce_once<M>(r);
  return r;
}

// ---- r = a - b mod m  (a, b canonical)          field.rs:217-240 / scalar.rs:210-237 -------------------
template <class M>
__device__ __forceinline__ Fe fe_sub(const Fe& a, const Fe& b) {
  Fe r;
  uint32_t mask;
  asm("sub.cc.u32  %0, %9,  %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32    %8, 0, 0;\n\t"
      : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]), "=r"(mask)
      : "r"(a.w[0]), "r"(a.w[1]), "r"(a.w[2]), "r"(a.w[3]), "r"(a.w[4]), "r"(a.w[5]), "r"(a.w[6]), "r"(a.w[7]),
        "r"(b.w[0]), "r"(b.w[1]), "r"(b.w[2]), "r"(b.w[3]), "r"(b.w[4]), "r"(b.w[5]), "r"(b.w[6]), "r"(b.w[7]));
  // borrow -> add the modulus back (the reference's underflow_mask, field.rs:232)
  asm("add.cc.u32  %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, %9;\n\t"
      "addc.cc.u32 %2, %2, %10;\n\t"
      "addc.cc.u32 %3, %3, %11;\n\t"
      "addc.cc.u32 %4, %4, 0;\n\t"
      "addc.cc.u32 %5, %5, 0;\n\t"
      "addc.cc.u32 %6, %6, 0;\n\t"
      "addc.u32    %7, %7, %12;\n\t"
      : "+r"(r.w[0]), "+r"(r.w[1]), "+r"(r.w[2]), "+r"(r.w[3]), "+r"(r.w[4]), "+r"(r.w[5]), "+r"(r.w[6]), "+r"(r.w[7])
      : "r"(mask & M::M0), "r"(mask & M::M1), "r"(mask & M::M2), "r"(mask & M::M3), "r"(mask & M::M7));
  return r;
}

template <class M>
__device__ __forceinline__ Fe fe_neg(const Fe& a) {   // field.rs:170-189 (0 - a)
  Fe z{{0, 0, 0, 0, 0, 0, 0, 0}};
  return fe_sub<M>(z, a);
}

// ---- one CIOS row:  t = (t + a*bi + mq*m) / 2^32 -------------------------------------------------------
// t has 9 live words on entry (t8 == 0) and 8 on exit (the caller renames t1..t8 -> t0..t7).
template <class M>
__device__ __forceinline__ void cios_row(uint32_t (&t)[9], const Fe& a, uint32_t bi) {
  asm("{\n\t"
      ".reg .u32 mq, lo, hi, z;\n\t"
      // even words of a
      "mad.lo.cc.u32   %0, %9,  %17, %0;\n\t"
      "madc.hi.cc.u32  %1, %9,  %17, %1;\n\t"
      "madc.lo.cc.u32  %2, %11, %17, %2;\n\t"
      "madc.hi.cc.u32  %3, %11, %17, %3;\n\t"
      "madc.lo.cc.u32  %4, %13, %17, %4;\n\t"
      "madc.hi.cc.u32  %5, %13, %17, %5;\n\t"
      "madc.lo.cc.u32  %6, %15, %17, %6;\n\t"
      "madc.hi.cc.u32  %7, %15, %17, %7;\n\t"
      "addc.u32        %8, %8, 0;\n\t"
      // odd words of a
      "mad.lo.cc.u32   %1, %10, %17, %1;\n\t"
      "madc.hi.cc.u32  %2, %10, %17, %2;\n\t"
      "madc.lo.cc.u32  %3, %12, %17, %3;\n\t"
      "madc.hi.cc.u32  %4, %12, %17, %4;\n\t"
      "madc.lo.cc.u32  %5, %14, %17, %5;\n\t"
      "madc.hi.cc.u32  %6, %14, %17, %6;\n\t"
      "madc.lo.cc.u32  %7, %16, %17, %7;\n\t"
      "madc.hi.u32     %8, %16, %17, %8;\n\t"
      // Montgomery quotient digit
      "mul.lo.u32      mq, %0, %18;\n\t"
      "shl.b32         lo, mq, %23;\n\t"
      "shr.b32         hi, mq, %24;\n\t"
      // + mq * m, even words (m4..m6 are zero, m7 is one bit -> shifts)
      "mad.lo.cc.u32   z,  mq, %19, %0;\n\t"
      "madc.hi.cc.u32  %1, mq, %19, %1;\n\t"
      "madc.lo.cc.u32  %2, mq, %21, %2;\n\t"
      "madc.hi.cc.u32  %3, mq, %21, %3;\n\t"
      "addc.cc.u32     %4, %4, 0;\n\t"
      "addc.cc.u32     %5, %5, 0;\n\t"
      "addc.cc.u32     %6, %6, 0;\n\t"
      "addc.cc.u32     %7, %7, lo;\n\t"
      "addc.u32        %8, %8, hi;\n\t"
      // + mq * m, odd words
      "mad.lo.cc.u32   %1, mq, %20, %1;\n\t"
      "madc.hi.cc.u32  %2, mq, %20, %2;\n\t"
      "madc.lo.cc.u32  %3, mq, %22, %3;\n\t"
      "madc.hi.cc.u32  %4, mq, %22, %4;\n\t"
      "addc.cc.u32     %5, %5, 0;\n\t"
      "addc.cc.u32     %6, %6, 0;\n\t"
      "addc.cc.u32     %7, %7, 0;\n\t"
      "addc.u32        %8, %8, 0;\n\t"
      "}\n\t"
      : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8])
      : "r"(a.w[0]), "r"(a.w[1]), "r"(a.w[2]), "r"(a.w[3]), "r"(a.w[4]), "r"(a.w[5]), "r"(a.w[6]), "r"(a.w[7]),
        "r"(bi), "r"(M::NINV), "r"(M::M0), "r"(M::M1), "r"(M::M2), "r"(M::M3), "n"(M::TOP), "n"(32 - M::TOP));
  t[0] = t[1]; t[1] = t[2]; t[2] = t[3]; t[3] = t[4]; t[4] = t[5]; t[5] = t[6]; t[6] = t[7]; t[7] = t[8]; t[8] = 0;
}

// ---- Montgomery product without the final subtraction: returns a*b/R + (< m), i.e. < 2m for a*b < R*m ----
template <class M>
__device__ __forceinline__ Fe mont_mul_lazy(const Fe& a, const Fe& b) {
  uint32_t t[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 8; i++) cios_row<M>(t, a, b.w[i]);
  Fe r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.w[i] = t[i];
  return r;
}

// ---- r = a*b/R mod m, canonical for canonical inputs ------------------------------------------------
template <class M>
__device__ __forceinline__ Fe mont_mul(const Fe& a, const Fe& b) {
  Fe r = mont_mul_lazy<M>(a, b);
  reduce_once<M>(r);
  return r;
}
template <class M>
__device__ __forceinline__ Fe mont_sqr(const Fe& a) { return mont_mul<M>(a, a); }

template <class M>
__device__ __forceinline__ Fe to_mont(const Fe& a) { return mont_mul<M>(a, Consts<M>::R2()); }

// a/R mod m: Montgomery reduction of a zero-extended residue (the reference's from_montgomery, field.rs:830-836)
template <class M>
__device__ __forceinline__ Fe from_mont(const Fe& a) {
  Fe one{{1, 0, 0, 0, 0, 0, 0, 0}};
  return mont_mul<M>(a, one);
}

// ---- radix-2^52 limbs (the reference's [u64;5], field.rs:31-32) <-> 8 x u32 --------------------------
__device__ __forceinline__ Fe fe_from_limbs52(uint64_t l0, uint64_t l1, uint64_t l2, uint64_t l3, uint64_t l4) {
  uint64_t w0 = l0 | (l1 << 52);
  uint64_t w1 = (l1 >> 12) | (l2 << 40);
  uint64_t w2 = (l2 >> 24) | (l3 << 28);
  uint64_t w3 = (l3 >> 36) | (l4 << 16);
  Fe r;
  r.w[0] = (uint32_t)w0; r.w[1] = (uint32_t)(w0 >> 32);
  r.w[2] = (uint32_t)w1; r.w[3] = (uint32_t)(w1 >> 32);
  r.w[4] = (uint32_t)w2; r.w[5] = (uint32_t)(w2 >> 32);
  r.w[6] = (uint32_t)w3; r.w[7] = (uint32_t)(w3 >> 32);
  return r;
}
__device__ __forceinline__ void fe_to_limbs52(const Fe& a, uint64_t (&l)[5]) {
  const uint64_t MASK = (1ull << 52) - 1;
  uint64_t w0 = a.w[0] | ((uint64_t)a.w[1] << 32);
  uint64_t w1 = a.w[2] | ((uint64_t)a.w[3] << 32);
  uint64_t w2 = a.w[4] | ((uint64_t)a.w[5] << 32);
  uint64_t w3 = a.w[6] | ((uint64_t)a.w[7] << 32);
  l[0] = w0 & MASK;
  l[1] = ((w0 >> 52) | (w1 << 12)) & MASK;
  l[2] = ((w1 >> 40) | (w2 << 24)) & MASK;
  l[3] = ((w2 >> 28) | (w3 << 36)) & MASK;
  l[4] = w3 >> 16;
}
__device__ __forceinline__ Fe fe_load52(const uint64_t* __restrict__ p) {
  return fe_from_limbs52(p[0], p[1], p[2], p[3], p[4]);
}
__device__ __forceinline__ void fe_store52(uint64_t* __restrict__ p, const Fe& a) {
  uint64_t l[5];
  fe_to_limbs52(a, l);
  p[0] = l[0]; p[1] = l[1]; p[2] = l[2]; p[3] = l[3]; p[4] = l[4];
}

__device__ __forceinline__ bool fe_is_zero(const Fe& a) {
  return (a.w[0] | a.w[1] | a.w[2] | a.w[3] | a.w[4] | a.w[5] | a.w[6] | a.w[7]) == 0;
}
__device__ __forceinline__ bool fe_eq(const Fe& a, const Fe& b) {
  uint32_t d = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) d |= a.w[i] ^ b.w[i];
  return d == 0;
}

}  // namespace zc
