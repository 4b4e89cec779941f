// Micro-benchmark: issue rates of the integer-multiply and FP64 pipes on sm_100a (design input for zc_fe.cuh).
// Each kernel runs ITER iterations of NCHAIN independent dependent-chains per thread; reports lanes/clk/SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(uint64_t* out, uint32_t seed) {
  uint32_t a = threadIdx.x * 2654435761u + seed, b = a ^ 0x9e3779b9u;
  uint64_t c0 = a, c1 = b, c2 = a + 1, c3 = b + 1, c4 = a + 2, c5 = b + 2, c6 = a + 3, c7 = b + 3;
  double d0 = a, d1 = b, d2 = a + 1.0, d3 = b + 1.0, d4 = a + 2.0, d5 = b + 2.0, d6 = a + 3.0, d7 = b + 3.0;
  double da = 1.0000001, db = 1e-9;
  for (int i = 0; i < ITER; i++) {
    if (MODE == 0) {   // IMAD.WIDE.U32: 32x32+64
#define W(c) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c) : "r"(a), "r"(b));
      W(c0) W(c1) W(c2) W(c3) W(c4) W(c5) W(c6) W(c7)
    } else if (MODE == 1) {   // IMAD lo 32
      uint32_t *p = (uint32_t*)&c0;
#define L(c) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(*(uint32_t*)&c) : "r"(a), "r"(b));
      L(c0) L(c1) L(c2) L(c3) L(c4) L(c5) L(c6) L(c7)
    } else if (MODE == 2) {   // IMAD.HI
#define H(c) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(*(uint32_t*)&c) : "r"(a), "r"(b));
      H(c0) H(c1) H(c2) H(c3) H(c4) H(c5) H(c6) H(c7)
    } else if (MODE == 3) {   // DFMA
#define D(d) asm volatile("fma.rz.f64 %0, %1, %0, %2;" : "+d"(d) : "d"(da), "d"(db));
      D(d0) D(d1) D(d2) D(d3) D(d4) D(d5) D(d6) D(d7)
    } else if (MODE == 4) {   // DFMA + IMAD.WIDE interleaved (are the pipes independent?)
      W(c0) D(d0) W(c1) D(d1) W(c2) D(d2) W(c3) D(d3) W(c4) D(d4) W(c5) D(d5) W(c6) D(d6) W(c7) D(d7)
    } else if (MODE == 5) {   // carry chain: mad.lo.cc / madc.hi.cc pairs as in cios_row
      uint32_t *q = (uint32_t*)&c0;
      asm volatile("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\t"
                   "madc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\t"
                   "madc.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.cc.u32 %5, %8, %9, %5;\n\t"
                   "madc.lo.cc.u32 %6, %8, %10, %6;\n\tmadc.hi.u32 %7, %8, %10, %7;\n\t"
                   : "+r"(q[0]), "+r"(q[1]), "+r"(q[2]), "+r"(q[3]), "+r"(q[4]), "+r"(q[5]), "+r"(q[6]), "+r"(q[7])
                   : "r"(a), "r"(b), "r"(a ^ b));
      uint32_t *q2 = (uint32_t*)&c4;
      asm volatile("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\t"
                   "madc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\t"
                   "madc.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.cc.u32 %5, %8, %9, %5;\n\t"
                   "madc.lo.cc.u32 %6, %8, %10, %6;\n\tmadc.hi.u32 %7, %8, %10, %7;\n\t"
                   : "+r"(q2[0]), "+r"(q2[1]), "+r"(q2[2]), "+r"(q2[3]), "+r"(q2[4]), "+r"(q2[5]), "+r"(q2[6]), "+r"(q2[7])
                   : "r"(b), "r"(a), "r"(a ^ b));
    } else if (MODE == 6) {  // IADD3 64-bit adds (alu pipe) alongside IMAD.WIDE
      W(c0) W(c1) W(c2) W(c3)
      c4 += c0; c5 += c1; c6 += c2; c7 += c3;
    } else if (MODE == 7) {  // IMAD.WIDE + FFMA interleaved? skip; plain 64-bit adds only
      asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(((uint32_t*)&c0)[0]), "+r"(((uint32_t*)&c0)[1]) : "r"(a), "r"(b));
      asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(((uint32_t*)&c1)[0]), "+r"(((uint32_t*)&c1)[1]) : "r"(a), "r"(b));
      asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(((uint32_t*)&c2)[0]), "+r"(((uint32_t*)&c2)[1]) : "r"(a), "r"(b));
      asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(((uint32_t*)&c3)[0]), "+r"(((uint32_t*)&c3)[1]) : "r"(a), "r"(b));
      asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(((uint32_t*)&c4)[0]), "+r"(((uint32_t*)&c4)[1]) : "r"(a), "r"(b));
      asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(((uint32_t*)&c5)[0]), "+r"(((uint32_t*)&c5)[1]) : "r"(a), "r"(b));
      asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(((uint32_t*)&c6)[0]), "+r"(((uint32_t*)&c6)[1]) : "r"(a), "r"(b));
      asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(((uint32_t*)&c7)[0]), "+r"(((uint32_t*)&c7)[1]) : "r"(a), "r"(b));
    }
  }
  uint64_t r = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
  double dr = d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r ^ (uint64_t)__double_as_longlong(dr);
}
template <int MODE> void run(const char* name, double ops_per_iter) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  uint64_t* out; cudaMalloc(&out, sms * 8 * 256 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms * 8, 256>>>(out, 1); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; r++) k<MODE><<<sms * 8, 256>>>(out, r);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  double total = (double)sms * 8 * 256 * ITER * ops_per_iter;
  printf("%-28s %8.3f ms  %8.2f Gops/s  %6.2f lanes/clk/SM (at %d MHz nominal)\n", name, ms, total / ms / 1e6, total / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
  cudaFree(out);
}
int main() {
  run<0>("IMAD.WIDE.U32", 8);
  run<1>("IMAD lo", 8);
  run<2>("IMAD.HI", 8);
  run<3>("DFMA", 8);
  run<4>("DFMA+IMAD.WIDE (16 ops)", 16);
  run<5>("mad.lo.cc/madc.hi.cc (16)", 16);
  run<6>("4 IMAD.WIDE + 4 add64", 8);
  run<7>("add64 (cc pairs)", 8);
  return 0;
}
