// Single-warp latency of a carry-flag-free Montgomery multiplier written in plain C++ (64-bit temporaries), which ptxas
// is free to interleave across independent products -- unlike the mad.cc / madc chains of zc_fe.cuh.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../dusk_zerocaf_b200/csrc/zc_fe.cuh"
using namespace zc;
template <class M>
__device__ __forceinline__ Fe mont_mul_c(const Fe& a, const Fe& b) {
  const uint32_t m[4] = {M::M0, M::M1, M::M2, M::M3};
  uint32_t t[10];
#pragma unroll
  for (int k = 0; k < 10; k++) t[k] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint64_t c = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      uint64_t x = (uint64_t)a.w[j] * b.w[i] + t[j] + c;
      t[j] = (uint32_t)x; c = x >> 32;
    }
    uint64_t x = (uint64_t)t[8] + c; t[8] = (uint32_t)x; t[9] = (uint32_t)(x >> 32);
    const uint32_t q = t[0] * M::NINV;
    c = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      uint64_t y = (uint64_t)q * m[j] + t[j] + c;
      if (j > 0) t[j - 1] = (uint32_t)y;
      c = y >> 32;
    }
#pragma unroll
    for (int j = 4; j < 7; j++) { uint64_t y = (uint64_t)t[j] + c; t[j - 1] = (uint32_t)y; c = y >> 32; }
    { uint64_t y = (uint64_t)t[7] + (uint64_t)(q << M::TOP) + c; t[6] = (uint32_t)y; c = y >> 32; }
    { uint64_t y = (uint64_t)t[8] + (uint64_t)(q >> (32 - M::TOP)) + c; t[7] = (uint32_t)y; c = y >> 32; }
    t[8] = t[9] + (uint32_t)c; t[9] = 0;
  }
  Fe r;
#pragma unroll
  for (int k = 0; k < 8; k++) r.w[k] = t[k];
  reduce_once<M>(r);
  return r;
}
template <int K, int V>
__global__ void __launch_bounds__(32) chainK(const uint32_t* in, uint32_t* out, int iters) {
  Fe x[K], y;
#pragma unroll
  for (int j = 0; j < K; j++)
#pragma unroll
    for (int k = 0; k < 8; k++) x[j].w[k] = in[64 * threadIdx.x + 8 * j + k] & (k == 7 ? 0x0fffffffu : 0xffffffffu);
#pragma unroll
  for (int k = 0; k < 8; k++) y.w[k] = in[64 * threadIdx.x + 56 + k] & (k == 7 ? 0x0fffffffu : 0xffffffffu);
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < K; j++) x[j] = V ? mont_mul_c<ModP>(x[j], y) : mont_mul<ModP>(x[j], y);
  }
#pragma unroll
  for (int j = 0; j < K; j++)
#pragma unroll
    for (int k = 0; k < 8; k++) out[64 * threadIdx.x + 8 * j + k] = x[j].w[k];
}
int main() {
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  uint32_t *in, *o0, *o1; cudaMalloc(&in, 32 * 256); cudaMalloc(&o0, 32 * 256); cudaMalloc(&o1, 32 * 256);
  uint32_t h[2048]; for (int k = 0; k < 2048; k++) h[k] = 0x9e3779b9u * (k + 1) ^ (k << 7);
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 1000;
  chainK<4, 0><<<1, 32>>>(in, o0, 5); chainK<4, 1><<<1, 32>>>(in, o1, 5);
  uint32_t a[2048], b[2048]; cudaMemcpy(a, o0, sizeof(a), cudaMemcpyDeviceToHost); cudaMemcpy(b, o1, sizeof(b), cudaMemcpyDeviceToHost);
  int bad = 0; for (int k = 0; k < 2048; k++) if ((k & 63) < 32) bad += a[k] != b[k];
  printf("mismatching words: %d\n", bad);
  for (int v = 0; v < 6; v++) {
    float ms = 0;
    for (int r = 0; r < 3; r++) {
      cudaEventRecord(e0);
      switch (v) {
        case 0: chainK<1, 0><<<1, 32>>>(in, o0, iters); break; case 1: chainK<4, 0><<<1, 32>>>(in, o0, iters); break;
        case 2: chainK<1, 1><<<1, 32>>>(in, o0, iters); break; case 3: chainK<2, 1><<<1, 32>>>(in, o0, iters); break;
        case 4: chainK<4, 1><<<1, 32>>>(in, o0, iters); break; case 5: chainK<6, 1><<<1, 32>>>(in, o0, iters); break;
      }
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
    }
    const char* nm[] = {"asm  x1", "asm  x4", "C++  x1", "C++  x2", "C++  x4", "C++  x6"};
    printf("%-8s %.0f ns per step (%.0f cycles)\n", nm[v], ms * 1e6 / iters, ms * 1e-3 * clk * 1e3 / iters);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
