// Micro-benchmark 5 (round 2): throughput of the a*b part of a 253-bit product in two shapes, full occupancy.
//   A. mul_wide_8x8 of zc_fe.cuh: 8 x 32-bit words, 64 wide multiplies in carry chains (IMAD.WIDE.U32[.X]).
//   B. 9 x 28-bit limbs, 81 plain mad.wide.u32 into 17 independent 64-bit column accumulators (no carry flags; a column
//      holds at most 9 products < 2^56), then a carry pass that brings the columns back to 28 bits.
// Question: does the carry-free shape issue enough faster (plain IMAD.WIDE is 2 cycles / warp-instruction / SMSP in
// pipes.cu, the chained form ~4.4) to pay for 27 % more multiplies and the carry pass?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o pipes5 pipes5.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../dusk_zerocaf_b200/csrc/zc_fe.cuh"
using namespace zc;
#define ITER 512
template <int MODE>
__global__ void __launch_bounds__(128, 4) k(uint32_t* out, const uint32_t* in) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (MODE == 0) {
    uint32_t a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { a[j] = in[(tid + 37 * j) & 4095]; b[j] = in[(tid + 91 * j + 5) & 4095]; }
#pragma unroll 1
    for (int i = 0; i < ITER; i++) {
      uint32_t t[16];
      mul_wide_8x8(t, a, b);
#pragma unroll
      for (int j = 0; j < 8; j++) a[j] = t[j] ^ t[j + 8];
    }
    uint32_t r = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) r ^= a[j];
    out[tid] = r;
  } else {
    uint32_t a[9], b[9];
#pragma unroll
    for (int j = 0; j < 9; j++) { a[j] = in[(tid + 37 * j) & 4095] & 0x0fffffffu; b[j] = in[(tid + 91 * j + 5) & 4095] & 0x0fffffffu; }
#pragma unroll 1
    for (int i = 0; i < ITER; i++) {
      uint64_t col[17];
#pragma unroll
      for (int kk = 0; kk < 17; kk++) col[kk] = 0;
#pragma unroll
      for (int x = 0; x < 9; x++)
#pragma unroll
        for (int y = 0; y < 9; y++) col[x + y] += (uint64_t)a[x] * b[y];
      if (MODE == 2) {          // + the carry pass: limbs back to 28 bits (18 limbs)
        uint64_t c = 0;
        uint32_t l[18];
#pragma unroll
        for (int kk = 0; kk < 17; kk++) { const uint64_t v = col[kk] + c; l[kk] = (uint32_t)v & 0x0fffffffu; c = v >> 28; }
        l[17] = (uint32_t)c;
#pragma unroll
        for (int j = 0; j < 9; j++) a[j] = (l[j] ^ l[j + 9]) & 0x0fffffffu;
      } else {
#pragma unroll
        for (int j = 0; j < 8; j++) a[j] = ((uint32_t)col[j] ^ (uint32_t)(col[j + 9] >> 7)) & 0x0fffffffu;
        a[8] = ((uint32_t)col[8] ^ (uint32_t)(col[16] >> 3)) & 0x0fffffffu;
      }
    }
    uint32_t r = 0;
#pragma unroll
    for (int j = 0; j < 9; j++) r ^= a[j];
    out[tid] = r;
  }
}
template <int MODE> void run(const char* name) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int blocks = sms * 16;
  uint32_t *out, *in; cudaMalloc(&out, blocks * 128 * 4); cudaMalloc(&in, 4096 * 4);
  uint32_t h[4096]; for (int i = 0; i < 4096; i++) h[i] = 0x9e3779b9u * (i + 1) ^ (i << 11);
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, 128>>>(out, in); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; r++) k<MODE><<<blocks, 128>>>(out, in);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  const double prods = (double)blocks * 128 * ITER;
  printf("%-52s %8.3f ms  %.3e products/s  %6.1f cycles per warp-product per SMSP (at %d MHz)\n", name, ms, prods / (ms * 1e-3),
         (ms * 1e-3) * (clk * 1e3) * sms * 4 / (prods / 32), clk / 1000);
}
int main() {
  run<0>("A: 8x8 words, 64 chained wide multiplies");
  run<1>("B: 9x9 28-bit limbs, 81 plain mad.wide (no carry pass)");
  run<2>("B: 9x9 28-bit limbs, 81 plain mad.wide + carry pass");
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
