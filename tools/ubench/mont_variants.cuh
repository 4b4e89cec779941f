// Candidate Montgomery multipliers for zc_fe.cuh (experiment harness, not part of the product library).
#pragma once
#include "../../dusk_zerocaf_b200/csrc/zc_fe.cuh"
namespace zcx {
using namespace zc;

// ---- variant B: even/odd accumulator arrays; the 32-bit shift after each row is absorbed by swapping the roles of
// the two arrays and by 3-operand wide multiply-adds that read the accumulator two words further up.
// T = sum even[k] 2^(32k) + sum odd[k] 2^(32(k+1)).
template <class M>
__device__ __forceinline__ void row_first(uint32_t (&ev)[8], uint32_t (&od)[8], const Fe& a, uint32_t bi) {
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    asm("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=r"(ev[j]), "=r"(ev[j + 1]) : "r"(a.w[j]), "r"(bi));
    asm("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=r"(od[j]), "=r"(od[j + 1]) : "r"(a.w[j + 1]), "r"(bi));
  }
}
// generic row: on entry `ev` is the array whose pairs are word-aligned with the (already shifted) value and `od`
// is the previous row's aligned array: od[0] == 0 (reduced away), od[1] is a lone word at offset 0,
// od[2..7] sit at offsets 1..6.
template <class M>
__device__ __forceinline__ void row_next(uint32_t (&ev)[8], uint32_t (&od)[8], const Fe& a, uint32_t bi) {
  asm("{\n\t"
      "add.cc.u32      %0, %0, %9;\n\t"                     // lone word
      // od[j],od[j+1] = a[j+1]*bi + od[j+2],od[j+3]   (shift by two words for free)
      "madc.lo.cc.u32  %8,  %17, %24, %10;\n\t"
      "madc.hi.cc.u32  %9,  %17, %24, %11;\n\t"
      "madc.lo.cc.u32  %10, %19, %24, %12;\n\t"
      "madc.hi.cc.u32  %11, %19, %24, %13;\n\t"
      "madc.lo.cc.u32  %12, %21, %24, %14;\n\t"
      "madc.hi.cc.u32  %13, %21, %24, %15;\n\t"
      "madc.lo.cc.u32  %14, %23, %24, 0;\n\t"
      "madc.hi.u32     %15, %23, %24, 0;\n\t"
      // ev pairs += a[even]*bi
      "mad.lo.cc.u32   %0, %16, %24, %0;\n\t"
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
// reduction step: ev/od += mq * m with mq = ev[0] * NINV; afterwards ev[0] == 0.
template <class M>
__device__ __forceinline__ void row_redc(uint32_t (&ev)[8], uint32_t (&od)[8]) {
  asm("{\n\t"
      ".reg .u32 mq, lo, hi;\n\t"
      "mul.lo.u32      mq, %0, %16;\n\t"
      "shl.b32         lo, mq, %21;\n\t"
      "shr.b32         hi, mq, %22;\n\t"
      // odd-aligned: m1, m3, (m5 = 0), m7 = 1 << TOP
      "mad.lo.cc.u32   %8,  mq, %18, %8;\n\t"
      "madc.hi.cc.u32  %9,  mq, %18, %9;\n\t"
      "madc.lo.cc.u32  %10, mq, %20, %10;\n\t"
      "madc.hi.cc.u32  %11, mq, %20, %11;\n\t"
      "addc.cc.u32     %12, %12, 0;\n\t"
      "addc.cc.u32     %13, %13, 0;\n\t"
      "addc.cc.u32     %14, %14, lo;\n\t"
      "addc.u32        %15, %15, hi;\n\t"
      // even-aligned: m0, m2, (m4 = m6 = 0)
      "mad.lo.cc.u32   %0, mq, %17, %0;\n\t"
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
template <class M>
__device__ __forceinline__ Fe mont_mul_lazy_B(const Fe& a, const Fe& b) {
  uint32_t ev[8], od[8];
  row_first<M>(ev, od, a, b.w[0]);
  row_redc<M>(ev, od);
#pragma unroll
  for (int i = 1; i < 8; i += 2) {
    row_next<M>(od, ev, a, b.w[i]);
    row_redc<M>(od, ev);
    if (i + 1 < 8) {
      row_next<M>(ev, od, a, b.w[i + 1]);
      row_redc<M>(ev, od);
    }
  }
  // after 8 rows the last aligned array is `od` (rows 1,3,5,7 use od as aligned); result = ev(aligned-next) ...
  // Final: value/2^32 = od_aligned>>32 + ev.  Here: last row used (od as ev-role, ev as od-role):
  //   aligned array A = od (A[0] == 0), other array O = ev at offsets 1..8.   r[k] = O[k] + A[k+1], r[7] = O[7] + carry
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
template <class M>
__device__ __forceinline__ Fe mont_mul_B(const Fe& a, const Fe& b) {
  Fe r = mont_mul_lazy_B<M>(a, b);
  reduce_once<M>(r);
  return r;
}
}  // namespace zcx
