// Single-warp latency of a dependent Montgomery-multiplication chain: fully unrolled CIOS vs a rolled two-row loop.
// (Question: are the MSM tail kernels -- one warp, straight-line 30 KB point additions -- instruction-fetch bound?)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../dusk_zerocaf_b200/csrc/zc_fe.cuh"
using namespace zc;
template <class M>
__device__ __forceinline__ Fe mont_mul_rolled(const Fe& a, const Fe& b) {
  uint32_t ev[8], od[8];
  mont_row_first<M>(ev, od, a, b.w[0]);
  mont_row_redc<M>(ev, od);
  // row 1 separately so that the loop body handles (even row, odd row) pairs with fixed roles
  mont_row_next<M>(od, ev, a, b.w[1]);
  mont_row_redc<M>(od, ev);
  uint32_t bw[8];
#pragma unroll
  for (int k = 0; k < 8; k++) bw[k] = b.w[k];
#pragma unroll 1
  for (int i = 2; i < 8; i += 2) {
    uint32_t b0 = 0, b1 = 0;
#pragma unroll
    for (int k = 2; k < 8; k += 2) { if (k == i) { b0 = bw[k]; b1 = bw[k + 1]; } }
    mont_row_next<M>(ev, od, a, b0);
    mont_row_redc<M>(ev, od);
    mont_row_next<M>(od, ev, a, b1);
    mont_row_redc<M>(od, ev);
  }
  Fe r;
  asm("add.cc.u32  %0, %8,  %16;\n\taddc.cc.u32 %1, %9,  %17;\n\taddc.cc.u32 %2, %10, %18;\n\taddc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\taddc.cc.u32 %5, %13, %21;\n\taddc.cc.u32 %6, %14, %22;\n\taddc.u32    %7, %15, 0;\n\t"
      : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
      : "r"(ev[0]), "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]),
        "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]));
  reduce_once<M>(r);
  return r;
}
template <int V>
__global__ void __launch_bounds__(32) chain(const uint32_t* in, uint32_t* out, int iters) {
  Fe x, y;
#pragma unroll
  for (int k = 0; k < 8; k++) { x.w[k] = in[16 * threadIdx.x + k]; y.w[k] = in[16 * threadIdx.x + 8 + k]; }
  x.w[7] &= 0x0fffffffu; y.w[7] &= 0x0fffffffu;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    if (V == 0) { x = mont_mul<ModP>(x, y); y = mont_mul<ModP>(y, x); }
    else if (V == 1) { x = mont_mul_rolled<ModP>(x, y); y = mont_mul_rolled<ModP>(y, x); }
    else {   // 8 different unrolled multiplications in sequence (a 30 KB straight-line body, like one point addition)
      x = mont_mul<ModP>(x, y); y = mont_mul<ModP>(y, x); x = mont_mul<ModP>(x, y); y = mont_mul<ModP>(y, x);
      x = mont_mul<ModP>(x, y); y = mont_mul<ModP>(y, x); x = mont_mul<ModP>(x, y); y = mont_mul<ModP>(y, x);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; k++) { out[16 * threadIdx.x + k] = x.w[k]; out[16 * threadIdx.x + 8 + k] = y.w[k]; }
}
int main() {
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  uint32_t *in, *o; cudaMalloc(&in, 32 * 64); cudaMalloc(&o, 32 * 64 * 3);
  cudaMemset(in, 0x5b, 32 * 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  uint32_t h[3][512];
  for (int v = 0; v < 3; v++) {
    int iters = v == 2 ? 250 : 1000;
    for (int r = 0; r < 3; r++) {
      cudaEventRecord(e0);
      if (v == 0) chain<0><<<1, 32>>>(in, o, iters); else if (v == 1) chain<1><<<1, 32>>>(in, o + 512, iters); else chain<2><<<1, 32>>>(in, o + 1024, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (r == 2) printf("variant %d: %.1f ns per multiplication (%.0f cycles at %d MHz)\n", v, ms * 1e6 / 2000, ms * 1e-3 * clk * 1e3 / 2000, clk / 1000);
    }
    cudaMemcpy(h[v], o + 512 * v, 2048, cudaMemcpyDeviceToHost);
  }
  int bad = 0; for (int k = 0; k < 512; k++) bad += (h[0][k] != h[1][k]) + (h[0][k] != h[2][k]);
  printf("mismatches: %d\n", bad);
  return 0;
}
