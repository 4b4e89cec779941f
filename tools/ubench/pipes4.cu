// Micro-benchmark 4: row pattern  acc[j] += a[j] * b  (b shared -> operand reuse cache), with and without carry chain.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, const uint32_t* in) {
  uint32_t a[8], t[16], b = in[threadIdx.x + 500];
#pragma unroll
  for (int j = 0; j < 8; j++) { a[j] = in[threadIdx.x + 32 * j]; t[2 * j] = a[j] + 1; t[2 * j + 1] = a[j] * 3; }
#pragma unroll 1
  for (int i = 0; i < ITER; i++) {
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 8; j++)
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(t[2 * j]), "+r"(t[2 * j + 1]) : "r"(a[j]), "r"(b));
    } else if (MODE == 1) {
      asm volatile("mad.lo.cc.u32 %0, %16, %24, %0;\n\tmadc.hi.cc.u32 %1, %16, %24, %1;\n\t"
                   "madc.lo.cc.u32 %2, %17, %24, %2;\n\tmadc.hi.cc.u32 %3, %17, %24, %3;\n\t"
                   "madc.lo.cc.u32 %4, %18, %24, %4;\n\tmadc.hi.cc.u32 %5, %18, %24, %5;\n\t"
                   "madc.lo.cc.u32 %6, %19, %24, %6;\n\tmadc.hi.cc.u32 %7, %19, %24, %7;\n\t"
                   "madc.lo.cc.u32 %8, %20, %24, %8;\n\tmadc.hi.cc.u32 %9, %20, %24, %9;\n\t"
                   "madc.lo.cc.u32 %10, %21, %24, %10;\n\tmadc.hi.cc.u32 %11, %21, %24, %11;\n\t"
                   "madc.lo.cc.u32 %12, %22, %24, %12;\n\tmadc.hi.cc.u32 %13, %22, %24, %13;\n\t"
                   "madc.lo.cc.u32 %14, %23, %24, %14;\n\tmadc.hi.u32 %15, %23, %24, %15;\n\t"
                   : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]),
                     "+r"(t[8]), "+r"(t[9]), "+r"(t[10]), "+r"(t[11]), "+r"(t[12]), "+r"(t[13]), "+r"(t[14]), "+r"(t[15])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b));
    } else if (MODE == 2) {   // two independent 4-long chains
      asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\t"
                   "madc.lo.cc.u32 %2, %9, %12, %2;\n\tmadc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                   "madc.lo.cc.u32 %4, %10, %12, %4;\n\tmadc.hi.cc.u32 %5, %10, %12, %5;\n\t"
                   "madc.lo.cc.u32 %6, %11, %12, %6;\n\tmadc.hi.u32 %7, %11, %12, %7;\n\t"
                   : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b));
      asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\t"
                   "madc.lo.cc.u32 %2, %9, %12, %2;\n\tmadc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                   "madc.lo.cc.u32 %4, %10, %12, %4;\n\tmadc.hi.cc.u32 %5, %10, %12, %5;\n\t"
                   "madc.lo.cc.u32 %6, %11, %12, %6;\n\tmadc.hi.u32 %7, %11, %12, %7;\n\t"
                   : "+r"(t[8]), "+r"(t[9]), "+r"(t[10]), "+r"(t[11]), "+r"(t[12]), "+r"(t[13]), "+r"(t[14]), "+r"(t[15])
                   : "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b));
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int j = 0; j < 16; j++) r ^= t[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  uint32_t *out, *in; cudaMalloc(&out, sms * 8 * 256 * 4); cudaMalloc(&in, 8192); cudaMemset(in, 0x5a, 8192);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms * 8, 256>>>(out, in); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; r++) k<MODE><<<sms * 8, 256>>>(out, in);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  double total = (double)sms * 8 * 256 * ITER * 8;
  printf("%-44s %8.3f ms  %6.2f wide-mults/clk/SM  (%.2f cycles per warp-instr per SMSP)\n", name, ms, total / (ms * 1e-3) / sms / (clk * 1e3),
         128.0 / (total / (ms * 1e-3) / sms / (clk * 1e3)));
}
int main() {
  run<0>("8 x plain wide, shared b");
  run<1>("8-long carry chain, shared b");
  run<2>("2 x 4-long carry chains, shared b");
  return 0;
}
