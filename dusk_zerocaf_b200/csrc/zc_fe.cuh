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
// The Montgomery multiply is a CIOS loop, fully unrolled, written as PTX mad.lo.cc / madc.hi.cc carry chains that
// ptxas fuses into IMAD.WIDE.U32[.X]; two accumulator arrays (even- and odd-aligned pairs) keep every 64-bit
// product on an aligned register pair, so the loop has no register moves (measured 1.22x over a single-array CIOS).
// Products of NORMAL-form operands (the ABI's element-wise mul / square) skip Montgomery altogether: a full 8x8
// product followed by two folds with 2^K = -c (mod m), 112 wide multiplies instead of 2 x 96 + conversions.
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

// ---- Montgomery product, CIOS with split accumulators --------------------------------------------------------
// The running value is kept in two register arrays whose 64-bit pairs never straddle:
//     T = sum_k ev[k] 2^(32k)  +  sum_k od[k] 2^(32(k+1))
// so every 32x32+64 multiply-add lands on an aligned register pair (IMAD.WIDE.U32[.X] with no register moves).  The
// division by 2^32 after each row is absorbed by swapping the roles of the two arrays; the one realignment that swap
// needs is done by 3-operand multiply-adds that read their addend two words further up (mont_row_next).
template <class M>
__device__ __forceinline__ void mont_row_first(uint32_t (&ev)[8], uint32_t (&od)[8], const Fe& a, uint32_t bi) {
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    asm("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=r"(ev[j]), "=r"(ev[j + 1]) : "r"(a.w[j]), "r"(bi));
    asm("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=r"(od[j]), "=r"(od[j + 1]) : "r"(a.w[j + 1]), "r"(bi));
  }
}
// On entry `od` is the array that was word-aligned in the previous row: od[0] == 0 (reduced away), od[1] is a lone
// word at the new offset 0 and od[2..7] sit at offsets 1..6; `ev` is aligned with the shifted value.
template <class M>
__device__ __forceinline__ void mont_row_next(uint32_t (&ev)[8], uint32_t (&od)[8], const Fe& a, uint32_t bi) {
  asm("{\n\t"
      "add.cc.u32      %0, %0, %9;\n\t"                     // the lone word
      "madc.lo.cc.u32  %8,  %17, %24, %10;\n\t"             // od[j], od[j+1] = a[j+1]*bi + od[j+2], od[j+3]
      "madc.hi.cc.u32  %9,  %17, %24, %11;\n\t"
      "madc.lo.cc.u32  %10, %19, %24, %12;\n\t"
      "madc.hi.cc.u32  %11, %19, %24, %13;\n\t"
      "madc.lo.cc.u32  %12, %21, %24, %14;\n\t"
      "madc.hi.cc.u32  %13, %21, %24, %15;\n\t"
      "madc.lo.cc.u32  %14, %23, %24, 0;\n\t"
      "madc.hi.u32     %15, %23, %24, 0;\n\t"
      "mad.lo.cc.u32   %0, %16, %24, %0;\n\t"               // ev pairs += a[even] * bi
      "madc.hi.cc.u32  %1, %16, %24, %1;\n\t"
      "madc.lo.cc.u32  %2, %18, %24, %2;\n\t"
      "madc.hi.cc.u32  %3, %18, %24, %3;\n\t"
      "madc.lo.cc.u32  %4, %20, %24, %4;\n\t"
      "madc.hi.cc.u32  %5, %20, %24, %5;\n\t"
      "madc.lo.cc.u32  %6, %22, %24, %6;\n\t"
      "madc.hi.cc.u32  %7, %22, %24, %7;\n\t"
      "addc.u32        %15, %15, 0;\n\t"
      "}"
      : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7]),
        "+r"(od[0]), "+r"(od[1]), "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7])
      : "r"(a.w[0]), "r"(a.w[1]), "r"(a.w[2]), "r"(a.w[3]), "r"(a.w[4]), "r"(a.w[5]), "r"(a.w[6]), "r"(a.w[7]), "r"(bi));
}
// T += mq * m with mq = -T / m mod 2^32; afterwards ev[0] == 0.  Words 4..6 of m are zero and word 7 is one bit, so
// the step costs 4 wide multiplies, one low multiply and two shifts.
template <class M>
__device__ __forceinline__ void mont_row_redc(uint32_t (&ev)[8], uint32_t (&od)[8]) {
  asm("{\n\t"
      ".reg .u32 mq, lo, hi;\n\t"
      "mul.lo.u32      mq, %0, %16;\n\t"
      "shl.b32         lo, mq, %21;\n\t"
      "shr.b32         hi, mq, %22;\n\t"
      "mad.lo.cc.u32   %8,  mq, %18, %8;\n\t"               // odd-aligned words of m: m1, m3, (m5 = 0), m7 = 1 << TOP
      "madc.hi.cc.u32  %9,  mq, %18, %9;\n\t"
      "madc.lo.cc.u32  %10, mq, %20, %10;\n\t"
      "madc.hi.cc.u32  %11, mq, %20, %11;\n\t"
      "addc.cc.u32     %12, %12, 0;\n\t"
      "addc.cc.u32     %13, %13, 0;\n\t"
      "addc.cc.u32     %14, %14, lo;\n\t"
      "addc.u32        %15, %15, hi;\n\t"
      "mad.lo.cc.u32   %0, mq, %17, %0;\n\t"                // even-aligned words of m: m0, m2, (m4 = m6 = 0)
      "madc.hi.cc.u32  %1, mq, %17, %1;\n\t"
      "madc.lo.cc.u32  %2, mq, %19, %2;\n\t"
      "madc.hi.cc.u32  %3, mq, %19, %3;\n\t"
      "addc.cc.u32     %4, %4, 0;\n\t"
      "addc.cc.u32     %5, %5, 0;\n\t"
      "addc.cc.u32     %6, %6, 0;\n\t"
      "addc.cc.u32     %7, %7, 0;\n\t"
      "addc.u32        %15, %15, 0;\n\t"
      "}"
      : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7]),
        "+r"(od[0]), "+r"(od[1]), "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7])
      : "r"(M::NINV), "r"(M::M0), "r"(M::M1), "r"(M::M2), "r"(M::M3), "n"(M::TOP), "n"(32 - M::TOP));
}

// ---- Montgomery product without the final subtraction: returns a*b/R + (< m), i.e. < 2m for a*b < R*m ----
template <class M>
__device__ __forceinline__ Fe mont_mul_lazy(const Fe& a, const Fe& b) {
  uint32_t ev[8], od[8];
  mont_row_first<M>(ev, od, a, b.w[0]);
  mont_row_redc<M>(ev, od);
#pragma unroll
  for (int i = 1; i < 8; i += 2) {
    mont_row_next<M>(od, ev, a, b.w[i]);
    mont_row_redc<M>(od, ev);
    if (i + 1 < 8) {
      mont_row_next<M>(ev, od, a, b.w[i + 1]);
      mont_row_redc<M>(ev, od);
    }
  }
  // the last row ran with od aligned (od[0] == 0) and ev one word up: result word k = ev[k] + od[k+1]
  Fe r;
  asm("add.cc.u32  %0, %8,  %16;\n\t"
      "addc.cc.u32 %1, %9,  %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32    %7, %15, 0;\n\t"
      : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
      : "r"(ev[0]), "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]),
        "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]));
  return r;
}

// ---- r = a*b/R mod m, canonical for canonical inputs ------------------------------------------------
template <class M>
__device__ __forceinline__ Fe mont_mul(const Fe& a, const Fe& b) {
  Fe r = mont_mul_lazy<M>(a, b);
  reduce_once<M>(r);
  return r;
}
// ---- Montgomery squaring: a^2 / R mod m with 36 + 32 wide multiplies instead of 64 + 32 --------------------------------
//   a^2 = sum_i a_i 2^(32i) * ( a_i 2^(32i) + 2 sum_{j>i} a_j 2^(32j) )
// so row i of the CIOS loop multiplies by a_i the vector  u = (0, .., 0, a_i, a_{i+1} << 1, (2a)_{i+2}, .., (2a)_7)  -- the
// doubled high part of a, word by word (the bit a_i hands up to word i+1 of 2a belongs to the excluded low part, hence the
// plain shift at position i+1).  The rows below are mont_row_next with the multiplies of the i leading zero words removed:
// in the odd-aligned chain a zero word still moves its pair two words down and passes the carry on (addc pairs), in the
// even-aligned chain it costs nothing.  Same window bounds as a product with a multiplicand < 2^256 (2a < 2^256 for
// a < 2^255); result (a^2 + Q m) / R < a^2 / R + m, i.e. < 2m for a < 2m.
// (A separate full square -- sqr_wide_8 -- followed by eight reduction-only rows was measured first: bit-exact, 11 % faster
// for a lone warp but SLOWER at full occupancy, 521 against 468 cycles per warp-squaring: its ~180 carry-ripple additions
// run at 2 cycles each on the integer ALU pipe, which then binds instead of the multiplier pipe.  tools/ubench/sqrbench.cu)
// ---- generated by tools/gen_sqr_rows.py: rows 1..7 of the interleaved squaring (do not edit by hand) ----
__device__ __forceinline__ void sqr_row_1(uint32_t (&ev)[8], uint32_t (&od)[8], const Fe& u, uint32_t bi) {
  asm("{\n\t"
      "add.cc.u32      %0, %0, %9;\n\t"
      "madc.lo.cc.u32  %8, %16, %23, %10;\n\t"
      "madc.hi.cc.u32  %9, %16, %23, %11;\n\t"
      "madc.lo.cc.u32  %10, %18, %23, %12;\n\t"
      "madc.hi.cc.u32  %11, %18, %23, %13;\n\t"
      "madc.lo.cc.u32  %12, %20, %23, %14;\n\t"
      "madc.hi.cc.u32  %13, %20, %23, %15;\n\t"
      "madc.lo.cc.u32  %14, %22, %23, 0;\n\t"
      "madc.hi.u32  %15, %22, %23, 0;\n\t"
      "mad.lo.cc.u32   %2, %17, %23, %2;\n\t"
      "madc.hi.cc.u32  %3, %17, %23, %3;\n\t"
      "madc.lo.cc.u32  %4, %19, %23, %4;\n\t"
      "madc.hi.cc.u32  %5, %19, %23, %5;\n\t"
      "madc.lo.cc.u32  %6, %21, %23, %6;\n\t"
      "madc.hi.cc.u32  %7, %21, %23, %7;\n\t"
      "addc.u32        %15, %15, 0;\n\t"
      "}"
      : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7]),
        "+r"(od[0]), "+r"(od[1]), "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7])
      : "r"(u.w[1]), "r"(u.w[2]), "r"(u.w[3]), "r"(u.w[4]), "r"(u.w[5]), "r"(u.w[6]), "r"(u.w[7]), "r"(bi));
}

__device__ __forceinline__ void sqr_row_2(uint32_t (&ev)[8], uint32_t (&od)[8], const Fe& u, uint32_t bi) {
  asm("{\n\t"
      "add.cc.u32      %0, %0, %9;\n\t"
      "addc.cc.u32     %8, %10, 0;\n\t"
      "addc.cc.u32     %9, %11, 0;\n\t"
      "madc.lo.cc.u32  %10, %17, %22, %12;\n\t"
      "madc.hi.cc.u32  %11, %17, %22, %13;\n\t"
      "madc.lo.cc.u32  %12, %19, %22, %14;\n\t"
      "madc.hi.cc.u32  %13, %19, %22, %15;\n\t"
      "madc.lo.cc.u32  %14, %21, %22, 0;\n\t"
      "madc.hi.u32  %15, %21, %22, 0;\n\t"
      "mad.lo.cc.u32   %2, %16, %22, %2;\n\t"
      "madc.hi.cc.u32  %3, %16, %22, %3;\n\t"
      "madc.lo.cc.u32  %4, %18, %22, %4;\n\t"
      "madc.hi.cc.u32  %5, %18, %22, %5;\n\t"
      "madc.lo.cc.u32  %6, %20, %22, %6;\n\t"
      "madc.hi.cc.u32  %7, %20, %22, %7;\n\t"
      "addc.u32        %15, %15, 0;\n\t"
      "}"
      : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7]),
        "+r"(od[0]), "+r"(od[1]), "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7])
      : "r"(u.w[2]), "r"(u.w[3]), "r"(u.w[4]), "r"(u.w[5]), "r"(u.w[6]), "r"(u.w[7]), "r"(bi));
}

__device__ __forceinline__ void sqr_row_3(uint32_t (&ev)[8], uint32_t (&od)[8], const Fe& u, uint32_t bi) {
  asm("{\n\t"
      "add.cc.u32      %0, %0, %9;\n\t"
      "addc.cc.u32     %8, %10, 0;\n\t"
      "addc.cc.u32     %9, %11, 0;\n\t"
      "madc.lo.cc.u32  %10, %16, %21, %12;\n\t"
      "madc.hi.cc.u32  %11, %16, %21, %13;\n\t"
      "madc.lo.cc.u32  %12, %18, %21, %14;\n\t"
      "madc.hi.cc.u32  %13, %18, %21, %15;\n\t"
      "madc.lo.cc.u32  %14, %20, %21, 0;\n\t"
      "madc.hi.u32  %15, %20, %21, 0;\n\t"
      "mad.lo.cc.u32   %4, %17, %21, %4;\n\t"
      "madc.hi.cc.u32  %5, %17, %21, %5;\n\t"
      "madc.lo.cc.u32  %6, %19, %21, %6;\n\t"
      "madc.hi.cc.u32  %7, %19, %21, %7;\n\t"
      "addc.u32        %15, %15, 0;\n\t"
      "}"
      : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7]),
        "+r"(od[0]), "+r"(od[1]), "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7])
      : "r"(u.w[3]), "r"(u.w[4]), "r"(u.w[5]), "r"(u.w[6]), "r"(u.w[7]), "r"(bi));
}

__device__ __forceinline__ void sqr_row_4(uint32_t (&ev)[8], uint32_t (&od)[8], const Fe& u, uint32_t bi) {
  asm("{\n\t"
      "add.cc.u32      %0, %0, %9;\n\t"
      "addc.cc.u32     %8, %10, 0;\n\t"
      "addc.cc.u32     %9, %11, 0;\n\t"
      "addc.cc.u32     %10, %12, 0;\n\t"
      "addc.cc.u32     %11, %13, 0;\n\t"
      "madc.lo.cc.u32  %12, %17, %20, %14;\n\t"
      "madc.hi.cc.u32  %13, %17, %20, %15;\n\t"
      "madc.lo.cc.u32  %14, %19, %20, 0;\n\t"
      "madc.hi.u32  %15, %19, %20, 0;\n\t"
      "mad.lo.cc.u32   %4, %16, %20, %4;\n\t"
      "madc.hi.cc.u32  %5, %16, %20, %5;\n\t"
      "madc.lo.cc.u32  %6, %18, %20, %6;\n\t"
      "madc.hi.cc.u32  %7, %18, %20, %7;\n\t"
      "addc.u32        %15, %15, 0;\n\t"
      "}"
      : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7]),
        "+r"(od[0]), "+r"(od[1]), "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7])
      : "r"(u.w[4]), "r"(u.w[5]), "r"(u.w[6]), "r"(u.w[7]), "r"(bi));
}

__device__ __forceinline__ void sqr_row_5(uint32_t (&ev)[8], uint32_t (&od)[8], const Fe& u, uint32_t bi) {
  asm("{\n\t"
      "add.cc.u32      %0, %0, %9;\n\t"
      "addc.cc.u32     %8, %10, 0;\n\t"
      "addc.cc.u32     %9, %11, 0;\n\t"
      "addc.cc.u32     %10, %12, 0;\n\t"
      "addc.cc.u32     %11, %13, 0;\n\t"
      "madc.lo.cc.u32  %12, %16, %19, %14;\n\t"
      "madc.hi.cc.u32  %13, %16, %19, %15;\n\t"
      "madc.lo.cc.u32  %14, %18, %19, 0;\n\t"
      "madc.hi.u32  %15, %18, %19, 0;\n\t"
      "mad.lo.cc.u32   %6, %17, %19, %6;\n\t"
      "madc.hi.cc.u32  %7, %17, %19, %7;\n\t"
      "addc.u32        %15, %15, 0;\n\t"
      "}"
      : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7]),
        "+r"(od[0]), "+r"(od[1]), "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7])
      : "r"(u.w[5]), "r"(u.w[6]), "r"(u.w[7]), "r"(bi));
}

__device__ __forceinline__ void sqr_row_6(uint32_t (&ev)[8], uint32_t (&od)[8], const Fe& u, uint32_t bi) {
  asm("{\n\t"
      "add.cc.u32      %0, %0, %9;\n\t"
      "addc.cc.u32     %8, %10, 0;\n\t"
      "addc.cc.u32     %9, %11, 0;\n\t"
      "addc.cc.u32     %10, %12, 0;\n\t"
      "addc.cc.u32     %11, %13, 0;\n\t"
      "addc.cc.u32     %12, %14, 0;\n\t"
      "addc.cc.u32     %13, %15, 0;\n\t"
      "madc.lo.cc.u32  %14, %17, %18, 0;\n\t"
      "madc.hi.u32  %15, %17, %18, 0;\n\t"
      "mad.lo.cc.u32   %6, %16, %18, %6;\n\t"
      "madc.hi.cc.u32  %7, %16, %18, %7;\n\t"
      "addc.u32        %15, %15, 0;\n\t"
      "}"
      : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7]),
        "+r"(od[0]), "+r"(od[1]), "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7])
      : "r"(u.w[6]), "r"(u.w[7]), "r"(bi));
}

__device__ __forceinline__ void sqr_row_7(uint32_t (&ev)[8], uint32_t (&od)[8], const Fe& u, uint32_t bi) {
  asm("{\n\t"
      "add.cc.u32      %0, %0, %9;\n\t"
      "addc.cc.u32     %8, %10, 0;\n\t"
      "addc.cc.u32     %9, %11, 0;\n\t"
      "addc.cc.u32     %10, %12, 0;\n\t"
      "addc.cc.u32     %11, %13, 0;\n\t"
      "addc.cc.u32     %12, %14, 0;\n\t"
      "addc.cc.u32     %13, %15, 0;\n\t"
      "madc.lo.cc.u32  %14, %16, %17, 0;\n\t"
      "madc.hi.u32  %15, %16, %17, 0;\n\t"
      "}"
      : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7]),
        "+r"(od[0]), "+r"(od[1]), "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7])
      : "r"(u.w[7]), "r"(bi));
}


template <class M>
__device__ __forceinline__ Fe mont_sqr_lazy(const Fe& a) {
  Fe u;                                              // words of 2a (a < 2^255)
  u.w[0] = a.w[0] << 1;
#pragma unroll
  for (int k = 1; k < 8; k++) u.w[k] = (a.w[k] << 1) | (a.w[k - 1] >> 31);
  uint32_t ev[8], od[8];
  {
    Fe u0 = u; u0.w[0] = a.w[0]; u0.w[1] = a.w[1] << 1;
    mont_row_first<M>(ev, od, u0, a.w[0]);
  }
  mont_row_redc<M>(ev, od);
#define ZC_SQR_ROW(I, X, Y) { Fe ui = u; ui.w[I] = a.w[I]; if (I + 1 < 8) ui.w[(I + 1) & 7] = a.w[(I + 1) & 7] << 1; sqr_row_##I(X, Y, ui, a.w[I]); mont_row_redc<M>(X, Y); }
  ZC_SQR_ROW(1, od, ev) ZC_SQR_ROW(2, ev, od) ZC_SQR_ROW(3, od, ev) ZC_SQR_ROW(4, ev, od)
  ZC_SQR_ROW(5, od, ev) ZC_SQR_ROW(6, ev, od) ZC_SQR_ROW(7, od, ev)
#undef ZC_SQR_ROW
  // as in mont_mul_lazy: result word k = ev[k] + od[k+1]
  Fe r;
  asm("add.cc.u32  %0, %8,  %16;\n\t"
      "addc.cc.u32 %1, %9,  %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32    %7, %15, 0;\n\t"
      : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
      : "r"(ev[0]), "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]),
        "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]));
  return r;
}
// r = a^2 / R mod m, canonical for a < 2m
template <class M>
__device__ __forceinline__ Fe mont_sqr(const Fe& a) {
  Fe r = mont_sqr_lazy<M>(a);
  reduce_once<M>(r);
  return r;
}

template <class M>
__device__ __forceinline__ Fe to_mont(const Fe& a) { return mont_mul<M>(a, Consts<M>::R2()); }

// a/R mod m: Montgomery reduction of a zero-extended residue (the reference's from_montgomery, field.rs:830-836)
template <class M>
__device__ __forceinline__ Fe from_mont(const Fe& a) {
  Fe one{{1, 0, 0, 0, 0, 0, 0, 0}};
  return mont_mul<M>(a, one);
}

// ============================ products of NORMAL-form operands (no Montgomery) ============================
// m = 2^K + c with c < 2^125 (4 words = M0..M3), so 2^K = -c (mod m).  For canonical a, b:
//   T  = (a << SA)(b << SB) = 2^(256-K) a b                     SA + SB = 256 - K       64 wide multiplies
//        (the shifts are folded into the limb unpacking, fe_load52_shl)
//   H  = T >> 256 = floor(ab / 2^K),   Lo = (T mod 2^256) >> (256-K) = ab mod 2^K
//   U  = c H  (< 2^380),   Uh = U >> K (< 2^128),   Ul = U mod 2^K                       32 wide multiplies
//   V  = c Uh (< 2^253)                                                                   16 wide multiplies
//   ab = Lo - Ul + V  (mod m),  in (-2^K, 2^(K+1)):  one conditional +m, one conditional -m.
// This is the reference's Mul / Square (field.rs:250-262, 302-315; scalar.rs:247-283) as a value: canonical in,
// canonical out, bit-identical limbs after repacking.

// One row of a schoolbook product with split accumulators: X and Y are the two arrays (see mont_mul_lazy), X gets
// a1,a3,a5,a7 (its top pair is fresh), Y gets a0,a2,a4,a6; the carry out of Y lands in X's fresh top word.
#define ZC_WIDE_ROW8(X, XB, Y, YB, A, BI)                                                                          \
  asm("{\n\t"                                                                                                      \
      "mad.lo.cc.u32   %0, %17, %24, %0;\n\t"                                                                      \
      "madc.hi.cc.u32  %1, %17, %24, %1;\n\t"                                                                      \
      "madc.lo.cc.u32  %2, %19, %24, %2;\n\t"                                                                      \
      "madc.hi.cc.u32  %3, %19, %24, %3;\n\t"                                                                      \
      "madc.lo.cc.u32  %4, %21, %24, %4;\n\t"                                                                      \
      "madc.hi.cc.u32  %5, %21, %24, %5;\n\t"                                                                      \
      "madc.lo.cc.u32  %6, %23, %24, 0;\n\t"                                                                       \
      "madc.hi.u32     %7, %23, %24, 0;\n\t"                                                                       \
      "mad.lo.cc.u32   %8,  %16, %24, %8;\n\t"                                                                     \
      "madc.hi.cc.u32  %9,  %16, %24, %9;\n\t"                                                                     \
      "madc.lo.cc.u32  %10, %18, %24, %10;\n\t"                                                                    \
      "madc.hi.cc.u32  %11, %18, %24, %11;\n\t"                                                                    \
      "madc.lo.cc.u32  %12, %20, %24, %12;\n\t"                                                                    \
      "madc.hi.cc.u32  %13, %20, %24, %13;\n\t"                                                                    \
      "madc.lo.cc.u32  %14, %22, %24, %14;\n\t"                                                                    \
      "madc.hi.cc.u32  %15, %22, %24, %15;\n\t"                                                                    \
      "addc.u32        %7, %7, 0;\n\t"                                                                             \
      "}"                                                                                                          \
      : "+r"(X[XB]), "+r"(X[XB + 1]), "+r"(X[XB + 2]), "+r"(X[XB + 3]), "+r"(X[XB + 4]), "+r"(X[XB + 5]),          \
        "=&r"(X[XB + 6]), "=&r"(X[XB + 7]),                                                                        \
        "+r"(Y[YB]), "+r"(Y[YB + 1]), "+r"(Y[YB + 2]), "+r"(Y[YB + 3]), "+r"(Y[YB + 4]), "+r"(Y[YB + 5]),          \
        "+r"(Y[YB + 6]), "+r"(Y[YB + 7])                                                                           \
      : "r"(A[0]), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(A[4]), "r"(A[5]), "r"(A[6]), "r"(A[7]), "r"(BI))

#define ZC_WIDE_ROW8_FIRST(EV, OD, A, BI)                                                                          \
  do {                                                                                                             \
    _Pragma("unroll") for (int j_ = 0; j_ < 8; j_ += 2) {                                                          \
      asm("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=r"(EV[j_]), "=r"(EV[j_ + 1]) : "r"(A[j_]), "r"(BI));      \
      asm("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=r"(OD[j_]), "=r"(OD[j_ + 1]) : "r"(A[j_ + 1]), "r"(BI));  \
    }                                                                                                              \
  } while (0)

// t[0..15] = a * b  (8 x 8 words, 64 wide multiplies)
__device__ __forceinline__ void mul_wide_8x8(uint32_t (&t)[16], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
  uint32_t (&ev)[16] = t;
  uint32_t od[14];
  ZC_WIDE_ROW8_FIRST(ev, od, a, b[0]);
  ZC_WIDE_ROW8(ev, 2, od, 0, a, b[1]);
  ZC_WIDE_ROW8(od, 2, ev, 2, a, b[2]);
  ZC_WIDE_ROW8(ev, 4, od, 2, a, b[3]);
  ZC_WIDE_ROW8(od, 4, ev, 4, a, b[4]);
  ZC_WIDE_ROW8(ev, 6, od, 4, a, b[5]);
  ZC_WIDE_ROW8(od, 6, ev, 6, a, b[6]);
  ZC_WIDE_ROW8(ev, 8, od, 6, a, b[7]);
  // t = ev + (od << 32)
  asm("add.cc.u32  %0, %0, %15;\n\t"
      "addc.cc.u32 %1, %1, %16;\n\t"
      "addc.cc.u32 %2, %2, %17;\n\t"
      "addc.cc.u32 %3, %3, %18;\n\t"
      "addc.cc.u32 %4, %4, %19;\n\t"
      "addc.cc.u32 %5, %5, %20;\n\t"
      "addc.cc.u32 %6, %6, %21;\n\t"
      "addc.cc.u32 %7, %7, %22;\n\t"
      "addc.cc.u32 %8, %8, %23;\n\t"
      "addc.cc.u32 %9, %9, %24;\n\t"
      "addc.cc.u32 %10, %10, %25;\n\t"
      "addc.cc.u32 %11, %11, %26;\n\t"
      "addc.cc.u32 %12, %12, %27;\n\t"
      "addc.cc.u32 %13, %13, %28;\n\t"
      "addc.u32    %14, %14, 0;\n\t"
      : "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8]), "+r"(t[9]),
        "+r"(t[10]), "+r"(t[11]), "+r"(t[12]), "+r"(t[13]), "+r"(t[14]), "+r"(t[15])
      : "r"(od[0]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]), "r"(od[8]),
        "r"(od[9]), "r"(od[10]), "r"(od[11]), "r"(od[12]), "r"(od[13]));
}

// t[0..15] = a^2: 28 cross products (rows below, same split-accumulator scheme as mul_wide_8x8), doubled, plus the 8
// squares a_i^2 -- 36 wide multiplies instead of 64.
__device__ __forceinline__ void sqr_wide_8(uint32_t (&t)[16], const uint32_t (&a)[8]) {
  uint32_t ev[14], od[14];          // cross terms: ev[2..13] at offsets 2..13, od[k] at offset k+1
#define ZC_MULW(LO, HI, X, Y) asm("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=r"(LO), "=r"(HI) : "r"(X), "r"(Y))
  // row 0: a[0] * a[1..7], all fresh
  ZC_MULW(od[0], od[1], a[1], a[0]); ZC_MULW(ev[2], ev[3], a[2], a[0]); ZC_MULW(od[2], od[3], a[3], a[0]);
  ZC_MULW(ev[4], ev[5], a[4], a[0]); ZC_MULW(od[4], od[5], a[5], a[0]); ZC_MULW(ev[6], ev[7], a[6], a[0]);
  ZC_MULW(od[6], od[7], a[7], a[0]);
  // row 1: a[1] * a[2..7]
  asm("{\n\t"
      "mad.lo.cc.u32   %0, %13, %18, %0;\n\t"
      "madc.hi.cc.u32  %1, %13, %18, %1;\n\t"
      "madc.lo.cc.u32  %2, %15, %18, %2;\n\t"
      "madc.hi.cc.u32  %3, %15, %18, %3;\n\t"
      "madc.lo.cc.u32  %4, %17, %18, 0;\n\t"
      "madc.hi.u32     %5, %17, %18, 0;\n\t"
      "mad.lo.cc.u32   %6, %12, %18, %6;\n\t"
      "madc.hi.cc.u32  %7, %12, %18, %7;\n\t"
      "madc.lo.cc.u32  %8, %14, %18, %8;\n\t"
      "madc.hi.cc.u32  %9, %14, %18, %9;\n\t"
      "madc.lo.cc.u32  %10, %16, %18, %10;\n\t"
      "madc.hi.cc.u32  %11, %16, %18, %11;\n\t"
      "addc.u32        %5, %5, 0;\n\t"
      "}"
      : "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7]), "=&r"(ev[8]), "=&r"(ev[9]), "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7])
      : "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[1]));
  // row 2: a[2] * a[3..7]
  asm("{\n\t"
      "mad.lo.cc.u32   %0, %10, %15, %0;\n\t"
      "madc.hi.cc.u32  %1, %10, %15, %1;\n\t"
      "madc.lo.cc.u32  %2, %12, %15, %2;\n\t"
      "madc.hi.cc.u32  %3, %12, %15, %3;\n\t"
      "madc.lo.cc.u32  %4, %14, %15, 0;\n\t"
      "madc.hi.u32     %5, %14, %15, 0;\n\t"
      "mad.lo.cc.u32   %6, %11, %15, %6;\n\t"
      "madc.hi.cc.u32  %7, %11, %15, %7;\n\t"
      "madc.lo.cc.u32  %8, %13, %15, %8;\n\t"
      "madc.hi.cc.u32  %9, %13, %15, %9;\n\t"
      "addc.u32        %5, %5, 0;\n\t"
      "}"
      : "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7]), "=&r"(od[8]), "=&r"(od[9]), "+r"(ev[6]), "+r"(ev[7]), "+r"(ev[8]), "+r"(ev[9])
      : "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[2]));
  // row 3: a[3] * a[4..7]
  asm("{\n\t"
      "mad.lo.cc.u32   %0, %9, %12, %0;\n\t"
      "madc.hi.cc.u32  %1, %9, %12, %1;\n\t"
      "madc.lo.cc.u32  %2, %11, %12, 0;\n\t"
      "madc.hi.u32     %3, %11, %12, 0;\n\t"
      "mad.lo.cc.u32   %4, %8, %12, %4;\n\t"
      "madc.hi.cc.u32  %5, %8, %12, %5;\n\t"
      "madc.lo.cc.u32  %6, %10, %12, %6;\n\t"
      "madc.hi.cc.u32  %7, %10, %12, %7;\n\t"
      "addc.u32        %3, %3, 0;\n\t"
      "}"
      : "+r"(ev[8]), "+r"(ev[9]), "=&r"(ev[10]), "=&r"(ev[11]), "+r"(od[6]), "+r"(od[7]), "+r"(od[8]), "+r"(od[9])
      : "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[3]));
  // row 4: a[4] * a[5..7]
  asm("{\n\t"
      "mad.lo.cc.u32   %0, %6, %9, %0;\n\t"
      "madc.hi.cc.u32  %1, %6, %9, %1;\n\t"
      "madc.lo.cc.u32  %2, %8, %9, 0;\n\t"
      "madc.hi.u32     %3, %8, %9, 0;\n\t"
      "mad.lo.cc.u32   %4, %7, %9, %4;\n\t"
      "madc.hi.cc.u32  %5, %7, %9, %5;\n\t"
      "addc.u32        %3, %3, 0;\n\t"
      "}"
      : "+r"(od[8]), "+r"(od[9]), "=&r"(od[10]), "=&r"(od[11]), "+r"(ev[10]), "+r"(ev[11])
      : "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[4]));
  // row 5: a[5] * a[6..7]
  asm("{\n\t"
      "mad.lo.cc.u32   %0, %5, %6, 0;\n\t"
      "madc.hi.u32     %1, %5, %6, 0;\n\t"
      "mad.lo.cc.u32   %2, %4, %6, %2;\n\t"
      "madc.hi.cc.u32  %3, %4, %6, %3;\n\t"
      "addc.u32        %1, %1, 0;\n\t"
      "}"
      : "=&r"(ev[12]), "=&r"(ev[13]), "+r"(od[10]), "+r"(od[11])
      : "r"(a[6]), "r"(a[7]), "r"(a[5]));
  // row 6: a[6] * a[7..7]
  asm("{\n\t"
      "mad.lo.cc.u32   %0, %2, %3, 0;\n\t"
      "madc.hi.u32     %1, %2, %3, 0;\n\t"
      "}"
      : "=&r"(od[12]), "=&r"(od[13])
      : "r"(a[7]), "r"(a[6]));
  // the squares
#pragma unroll
  for (int i = 0; i < 8; i++) ZC_MULW(t[2 * i], t[2 * i + 1], a[i], a[i]);
#undef ZC_MULW
  // c = ev + (od << 32): c[1] = od[0], c[k] = ev[k] + od[k-1] (k = 2..13), c[14] = od[13] + carry   (reuses od[])
  asm("add.cc.u32  %0, %0, %13;\n\t"
      "addc.cc.u32 %1, %1, %14;\n\t"
      "addc.cc.u32 %2, %2, %15;\n\t"
      "addc.cc.u32 %3, %3, %16;\n\t"
      "addc.cc.u32 %4, %4, %17;\n\t"
      "addc.cc.u32 %5, %5, %18;\n\t"
      "addc.cc.u32 %6, %6, %19;\n\t"
      "addc.cc.u32 %7, %7, %20;\n\t"
      "addc.cc.u32 %8, %8, %21;\n\t"
      "addc.cc.u32 %9, %9, %22;\n\t"
      "addc.cc.u32 %10, %10, %23;\n\t"
      "addc.cc.u32 %11, %11, %24;\n\t"
      "addc.u32    %12, %12, 0;\n\t"
      : "+r"(od[1]), "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7]), "+r"(od[8]), "+r"(od[9]),
        "+r"(od[10]), "+r"(od[11]), "+r"(od[12]), "+r"(od[13])
      : "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]), "r"(ev[8]), "r"(ev[9]), "r"(ev[10]),
        "r"(ev[11]), "r"(ev[12]), "r"(ev[13]));
  // now c[k] = od[k-1] for k = 1..14.  t += 2 c
  uint32_t d[15];                   // d[k-1] = word k of 2c, k = 1..15
  d[0] = od[0] << 1;
#pragma unroll
  for (int k = 2; k <= 14; k++) d[k - 1] = __funnelshift_l(od[k - 2], od[k - 1], 1);
  d[14] = od[13] >> 31;
  asm("add.cc.u32  %0, %0, %15;\n\t"
      "addc.cc.u32 %1, %1, %16;\n\t"
      "addc.cc.u32 %2, %2, %17;\n\t"
      "addc.cc.u32 %3, %3, %18;\n\t"
      "addc.cc.u32 %4, %4, %19;\n\t"
      "addc.cc.u32 %5, %5, %20;\n\t"
      "addc.cc.u32 %6, %6, %21;\n\t"
      "addc.cc.u32 %7, %7, %22;\n\t"
      "addc.cc.u32 %8, %8, %23;\n\t"
      "addc.cc.u32 %9, %9, %24;\n\t"
      "addc.cc.u32 %10, %10, %25;\n\t"
      "addc.cc.u32 %11, %11, %26;\n\t"
      "addc.cc.u32 %12, %12, %27;\n\t"
      "addc.cc.u32 %13, %13, %28;\n\t"
      "addc.u32    %14, %14, %29;\n\t"
      : "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8]), "+r"(t[9]),
        "+r"(t[10]), "+r"(t[11]), "+r"(t[12]), "+r"(t[13]), "+r"(t[14]), "+r"(t[15])
      : "r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3]), "r"(d[4]), "r"(d[5]), "r"(d[6]), "r"(d[7]), "r"(d[8]), "r"(d[9]),
        "r"(d[10]), "r"(d[11]), "r"(d[12]), "r"(d[13]), "r"(d[14]));
}

// u[0..11] = c * h   (c = the modulus' four low words, h 8 words; 32 wide multiplies)
template <class M>
__device__ __forceinline__ void mul_c_8(uint32_t (&u)[12], const uint32_t (&h)[8]) {
  uint32_t (&ev)[12] = u;
  uint32_t od[10];
  const uint32_t c0 = M::M0, c1 = M::M1, c2 = M::M2, c3 = M::M3;
  ZC_WIDE_ROW8_FIRST(ev, od, h, c0);
  ZC_WIDE_ROW8(ev, 2, od, 0, h, c1);
  ZC_WIDE_ROW8(od, 2, ev, 2, h, c2);
  ZC_WIDE_ROW8(ev, 4, od, 2, h, c3);
  asm("add.cc.u32  %0, %0, %11;\n\t"
      "addc.cc.u32 %1, %1, %12;\n\t"
      "addc.cc.u32 %2, %2, %13;\n\t"
      "addc.cc.u32 %3, %3, %14;\n\t"
      "addc.cc.u32 %4, %4, %15;\n\t"
      "addc.cc.u32 %5, %5, %16;\n\t"
      "addc.cc.u32 %6, %6, %17;\n\t"
      "addc.cc.u32 %7, %7, %18;\n\t"
      "addc.cc.u32 %8, %8, %19;\n\t"
      "addc.cc.u32 %9, %9, %20;\n\t"
      "addc.u32    %10, %10, 0;\n\t"
      : "+r"(u[1]), "+r"(u[2]), "+r"(u[3]), "+r"(u[4]), "+r"(u[5]), "+r"(u[6]), "+r"(u[7]), "+r"(u[8]), "+r"(u[9]),
        "+r"(u[10]), "+r"(u[11])
      : "r"(od[0]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]), "r"(od[8]),
        "r"(od[9]));
}

// v[0..7] = c * g   (g 4 words; 16 wide multiplies).  Rows run over the words of g, the row operand is c.
template <class M>
__device__ __forceinline__ void mul_c_4(uint32_t (&v)[8], const uint32_t (&g)[4]) {
  uint32_t (&ev)[8] = v;
  uint32_t od[6];
  const uint32_t c0 = M::M0, c1 = M::M1, c2 = M::M2, c3 = M::M3;
  asm("mul.lo.u32 %0, %4, %8;\n\tmul.hi.u32 %1, %4, %8;\n\tmul.lo.u32 %2, %6, %8;\n\tmul.hi.u32 %3, %6, %8;"
      : "=&r"(ev[0]), "=&r"(ev[1]), "=&r"(ev[2]), "=&r"(ev[3]) : "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(g[0]));
  asm("mul.lo.u32 %0, %5, %8;\n\tmul.hi.u32 %1, %5, %8;\n\tmul.lo.u32 %2, %7, %8;\n\tmul.hi.u32 %3, %7, %8;"
      : "=&r"(od[0]), "=&r"(od[1]), "=&r"(od[2]), "=&r"(od[3]) : "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(g[0]));
#define ZC_WIDE_ROW4(X, XB, Y, YB, BI)                                                                             \
  asm("{\n\t"                                                                                                      \
      "mad.lo.cc.u32   %0, %9,  %12, %0;\n\t"                                                                      \
      "madc.hi.cc.u32  %1, %9,  %12, %1;\n\t"                                                                      \
      "madc.lo.cc.u32  %2, %11, %12, 0;\n\t"                                                                       \
      "madc.hi.u32     %3, %11, %12, 0;\n\t"                                                                       \
      "mad.lo.cc.u32   %4, %8,  %12, %4;\n\t"                                                                      \
      "madc.hi.cc.u32  %5, %8,  %12, %5;\n\t"                                                                      \
      "madc.lo.cc.u32  %6, %10, %12, %6;\n\t"                                                                      \
      "madc.hi.cc.u32  %7, %10, %12, %7;\n\t"                                                                      \
      "addc.u32        %3, %3, 0;\n\t"                                                                             \
      "}"                                                                                                          \
      : "+r"(X[XB]), "+r"(X[XB + 1]), "=&r"(X[XB + 2]), "=&r"(X[XB + 3]),                                          \
        "+r"(Y[YB]), "+r"(Y[YB + 1]), "+r"(Y[YB + 2]), "+r"(Y[YB + 3])                                             \
      : "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(BI))
  ZC_WIDE_ROW4(ev, 2, od, 0, g[1]);
  ZC_WIDE_ROW4(od, 2, ev, 2, g[2]);
  ZC_WIDE_ROW4(ev, 4, od, 2, g[3]);
#undef ZC_WIDE_ROW4
  asm("add.cc.u32  %0, %0, %7;\n\t"
      "addc.cc.u32 %1, %1, %8;\n\t"
      "addc.cc.u32 %2, %2, %9;\n\t"
      "addc.cc.u32 %3, %3, %10;\n\t"
      "addc.cc.u32 %4, %4, %11;\n\t"
      "addc.cc.u32 %5, %5, %12;\n\t"
      "addc.u32    %6, %6, 0;\n\t"
      : "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7])
      : "r"(od[0]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]));
}

// K = bit length of the power of two in m; SA / SB = the two operand pre-shifts
template <class M> struct Shape { static constexpr int K = 224 + M::TOP, S = 256 - K, SA = S / 2, SB = S - SA; };

// fold a 16-word product t = 2^S * x (x < m^2, S <= 256 - K) to x mod m, canonical
template <class M, int S>
__device__ __forceinline__ Fe fold_product(const uint32_t (&t)[16]) {
  constexpr int TOP = M::TOP, K = Shape<M>::K, HS = K + S - 224;     // H = t >> (K + S) starts HS bits into word 7
  uint32_t h[8], u[12], g[4], v[8];
  static_assert(HS >= 1 && HS <= 32, "product shift out of range");
#pragma unroll
  for (int k = 0; k < 8; k++) h[k] = (HS == 32) ? t[8 + k] : __funnelshift_r(t[7 + k], k + 8 < 16 ? t[8 + k] : 0u, HS);
  mul_c_8<M>(u, h);
  // Uh = u >> K  (K = 224 + TOP)
#pragma unroll
  for (int k = 0; k < 4; k++) g[k] = __funnelshift_r(u[7 + k], u[8 + k], TOP);
  mul_c_4<M>(v, g);
  // Lo = (t mod 2^(K+S)) >> S
  Fe r;
#pragma unroll
  for (int k = 0; k < 7; k++) r.w[k] = __funnelshift_r(t[k], t[k + 1], S);
  r.w[7] = (HS == 32) ? (t[7] >> S) : ((t[7] & ((1u << HS) - 1u)) >> S);
  const uint32_t ul7 = u[7] & ((1u << TOP) - 1u);
  uint32_t bw;
  // r = Lo + V - Ul ; bw = all-ones when the result is negative
  asm("add.cc.u32  %0, %0, %9;\n\t"
      "addc.cc.u32 %1, %1, %10;\n\t"
      "addc.cc.u32 %2, %2, %11;\n\t"
      "addc.cc.u32 %3, %3, %12;\n\t"
      "addc.cc.u32 %4, %4, %13;\n\t"
      "addc.cc.u32 %5, %5, %14;\n\t"
      "addc.cc.u32 %6, %6, %15;\n\t"
      "addc.u32    %7, %7, %16;\n\t"
      "sub.cc.u32  %0, %0, %17;\n\t"
      "subc.cc.u32 %1, %1, %18;\n\t"
      "subc.cc.u32 %2, %2, %19;\n\t"
      "subc.cc.u32 %3, %3, %20;\n\t"
      "subc.cc.u32 %4, %4, %21;\n\t"
      "subc.cc.u32 %5, %5, %22;\n\t"
      "subc.cc.u32 %6, %6, %23;\n\t"
      "subc.cc.u32 %7, %7, %24;\n\t"
      "subc.u32    %8, 0, 0;\n\t"
      : "+r"(r.w[0]), "+r"(r.w[1]), "+r"(r.w[2]), "+r"(r.w[3]), "+r"(r.w[4]), "+r"(r.w[5]), "+r"(r.w[6]), "+r"(r.w[7]), "=r"(bw)
      : "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(ul7));
  asm("add.cc.u32  %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, %9;\n\t"
      "addc.cc.u32 %2, %2, %10;\n\t"
      "addc.cc.u32 %3, %3, %11;\n\t"
      "addc.cc.u32 %4, %4, 0;\n\t"
      "addc.cc.u32 %5, %5, 0;\n\t"
      "addc.cc.u32 %6, %6, 0;\n\t"
      "addc.u32    %7, %7, %12;\n\t"
      : "+r"(r.w[0]), "+r"(r.w[1]), "+r"(r.w[2]), "+r"(r.w[3]), "+r"(r.w[4]), "+r"(r.w[5]), "+r"(r.w[6]), "+r"(r.w[7])
      : "r"(bw & M::M0), "r"(bw & M::M1), "r"(bw & M::M2), "r"(bw & M::M3), "r"(bw & M::M7));
  reduce_once<M>(r);
  return r;
}

// a * b mod m from operands already shifted left by SA and SB bits (fe_load52_shl folds the shifts into the unpacking)
template <class M>
__device__ __forceinline__ Fe fe_mul_normal_pre(const Fe& a_shl, const Fe& b_shl) {
  uint32_t t[16];
  mul_wide_8x8(t, a_shl.w, b_shl.w);
  return fold_product<M, Shape<M>::S>(t);
}
// a^2 mod m from a << SA
template <class M>
__device__ __forceinline__ Fe fe_sqr_normal_pre(const Fe& a_shl) {
  uint32_t t[16];
  sqr_wide_8(t, a_shl.w);
  return fold_product<M, 2 * Shape<M>::SA>(t);
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
// value << S (S <= 4; the caller guarantees value << S < 2^256), shifts folded into the unpacking
template <int S>
__device__ __forceinline__ Fe fe_load52_shl(const uint64_t* __restrict__ p) {
  const uint64_t l0 = p[0], l1 = p[1], l2 = p[2], l3 = p[3], l4 = p[4];
  const uint64_t w0 = (l0 << S) | (l1 << (52 + S));
  const uint64_t w1 = (l1 >> (12 - S)) | (l2 << (40 + S));
  const uint64_t w2 = (l2 >> (24 - S)) | (l3 << (28 + S));
  const uint64_t w3 = (l3 >> (36 - S)) | (l4 << (16 + S));
  Fe r;
  r.w[0] = (uint32_t)w0; r.w[1] = (uint32_t)(w0 >> 32);
  r.w[2] = (uint32_t)w1; r.w[3] = (uint32_t)(w1 >> 32);
  r.w[4] = (uint32_t)w2; r.w[5] = (uint32_t)(w2 >> 32);
  r.w[6] = (uint32_t)w3; r.w[7] = (uint32_t)(w3 >> 32);
  return r;
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
