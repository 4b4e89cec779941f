// Micro-benchmark 3: IMAD.WIDE.U32 forms (plain / carry-out / .X carry-in) issue rate on sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 4096
#define NCH 8
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, const uint32_t* in) {
  uint32_t a[NCH], b[NCH], lo[NCH], hi[NCH], lo2[NCH], hi2[NCH];
#pragma unroll
  for (int j = 0; j < NCH; j++) { a[j] = in[threadIdx.x + 32 * j]; b[j] = in[threadIdx.x + 32 * j + 7]; lo[j] = a[j] + 1; hi[j] = b[j] + 2; lo2[j] = a[j] + 3; hi2[j] = b[j] + 5; }
#pragma unroll 1
  for (int i = 0; i < ITER; i++) {
#pragma unroll
    for (int j = 0; j < NCH; j++) {
      if (MODE == 0) {        // plain wide, no carry in/out
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[j]), "+r"(hi[j]) : "r"(a[j]), "r"(b[j]));
      } else if (MODE == 1) { // 2-long chain: wide with carry-out, then wide .X with carry-in
        asm volatile("mad.lo.cc.u32 %0, %4, %5, %0;\n\tmadc.hi.cc.u32 %1, %4, %5, %1;\n\t"
                     "madc.lo.cc.u32 %2, %4, %5, %2;\n\tmadc.hi.u32 %3, %4, %5, %3;"
                     : "+r"(lo[j]), "+r"(hi[j]), "+r"(lo2[j]), "+r"(hi2[j]) : "r"(a[j]), "r"(b[j]));
      } else if (MODE == 2) { // wide + separate 64-bit add of another pair (ALU) : carry handled on the ALU pipe
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[j]), "+r"(hi[j]) : "r"(a[j]), "r"(b[j]));
        asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo2[j]), "+r"(hi2[j]) : "r"(a[j]), "r"(b[j]));
      } else if (MODE == 3) { // mul.wide fresh (c = RZ) + 64-bit add on ALU
        uint32_t pl, ph;
        asm volatile("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=r"(pl), "=r"(ph) : "r"(a[j]), "r"(b[j]));
        asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo[j]), "+r"(hi[j]) : "r"(pl), "r"(ph));
      }
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int j = 0; j < NCH; j++) r ^= lo[j] ^ hi[j] ^ lo2[j] ^ hi2[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, double wide_per_iter) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  uint32_t *out, *in; cudaMalloc(&out, sms * 8 * 256 * 4); cudaMalloc(&in, 4096); cudaMemset(in, 0x5a, 4096);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms * 8, 256>>>(out, in); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; r++) k<MODE><<<sms * 8, 256>>>(out, in);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  double total = (double)sms * 8 * 256 * ITER * wide_per_iter;
  printf("%-44s %8.3f ms  %6.2f wide-mults/clk/SM (at %d MHz nominal)\n", name, ms, total / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
}
int main() {
  run<0>("IMAD.WIDE.U32 plain x8", NCH);
  run<1>("IMAD.WIDE.U32 P-out + IMAD.WIDE.U32.X x8", 2 * NCH);
  run<2>("IMAD.WIDE.U32 plain + add64 (ALU) x8", NCH);
  run<3>("mul.wide + add64 x8", NCH);
  return 0;
}
