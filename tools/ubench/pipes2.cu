// Micro-benchmark 2: IMAD.WIDE.U32 vs IMAD.WIDE.U32.X (carry-chained) issue rate on sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 2048
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed) {
  uint32_t a = threadIdx.x * 2654435761u + seed, b = a ^ 0x9e3779b9u, c = a + b;
  uint32_t t[16];
#pragma unroll
  for (int i = 0; i < 16; i++) t[i] = a + i;
#pragma unroll 1
  for (int i = 0; i < ITER; i++) {
    a ^= t[5]; b ^= t[9]; c ^= t[2];
    if (MODE == 0) {   // one 8-wide carry chain + second chain (16 wide mads, as cios_row's a*bi part)
      asm volatile("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\t"
                   "madc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\t"
                   "madc.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.cc.u32 %5, %8, %9, %5;\n\t"
                   "madc.lo.cc.u32 %6, %8, %10, %6;\n\tmadc.hi.u32 %7, %8, %10, %7;\n\t"
                   : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7])
                   : "r"(a), "r"(b), "r"(c));
      asm volatile("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\t"
                   "madc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\t"
                   "madc.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.cc.u32 %5, %8, %9, %5;\n\t"
                   "madc.lo.cc.u32 %6, %8, %10, %6;\n\tmadc.hi.u32 %7, %8, %10, %7;\n\t"
                   : "+r"(t[8]), "+r"(t[9]), "+r"(t[10]), "+r"(t[11]), "+r"(t[12]), "+r"(t[13]), "+r"(t[14]), "+r"(t[15])
                   : "r"(b), "r"(c), "r"(a));
    } else if (MODE == 1) {   // 8 independent mad.wide (no carry), 64-bit accumulators
#pragma unroll
      for (int j = 0; j < 8; j++) {
        uint64_t acc = ((uint64_t)t[2 * j + 1] << 32) | t[2 * j];
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(j & 1 ? b : c));
        t[2 * j] = (uint32_t)acc; t[2 * j + 1] = (uint32_t)(acc >> 32);
      }
    } else if (MODE == 2) {   // 8 mad.wide + 8 add.cc (64-bit adds of something) : typical "accumulate then carry-propagate"
#pragma unroll
      for (int j = 0; j < 8; j++) {
        uint64_t acc = ((uint64_t)t[2 * j + 1] << 32) | t[2 * j];
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(j & 1 ? b : c));
        t[2 * j] = (uint32_t)acc; t[2 * j + 1] = (uint32_t)(acc >> 32);
      }
      asm volatile("add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %9;\n\taddc.cc.u32 %2, %2, %8;\n\taddc.cc.u32 %3, %3, %9;\n\t"
                   "addc.cc.u32 %4, %4, %8;\n\taddc.cc.u32 %5, %5, %9;\n\taddc.cc.u32 %6, %6, %8;\n\taddc.u32 %7, %7, %9;\n\t"
                   : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7])
                   : "r"(t[8]), "r"(t[9]));
    } else if (MODE == 3) {   // separate mul.lo / mul.hi + add chains (no fused wide)
#pragma unroll
      for (int j = 0; j < 8; j++) {
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(t[2 * j]) : "r"(a), "r"(j & 1 ? b : c));
        asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(t[2 * j + 1]) : "r"(a), "r"(j & 1 ? b : c));
      }
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) r ^= t[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, double ops_per_iter) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  uint32_t* out; cudaMalloc(&out, sms * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms * 8, 256>>>(out, 1); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; r++) k<MODE><<<sms * 8, 256>>>(out, r);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  double total = (double)sms * 8 * 256 * ITER * ops_per_iter;
  printf("%-34s %8.3f ms  %6.2f wide-mults/clk/SM (at %d MHz nominal)\n", name, ms, total / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
  cudaFree(out);
}
int main() {
  run<0>("carry chain (8 wide per chain x2)", 8);
  run<1>("mad.wide no carry x8", 8);
  run<2>("mad.wide x8 + 8 addc", 8);
  run<3>("mad.lo + mad.hi separate x8", 8);
  return 0;
}
