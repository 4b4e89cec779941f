// Correctness + throughput of Montgomery-multiplier variants (dependent chains, registers only).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "mont_variants.cuh"
using namespace zc;
#define CHAIN 512
template <int V> __device__ __forceinline__ Fe mm(const Fe& a, const Fe& b) {
  if (V == 0) return mont_mul<ModP>(a, b);
  else return zcx::mont_mul_B<ModP>(a, b);
}
template <int V>
__global__ void __launch_bounds__(256) chain(const uint32_t* in, uint32_t* out, int iters) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  Fe x, y;
#pragma unroll
  for (int k = 0; k < 8; k++) { x.w[k] = in[16 * i + k]; y.w[k] = in[16 * i + 8 + k]; }
  x.w[7] &= 0x0fffffffu; y.w[7] &= 0x0fffffffu;
#pragma unroll 1
  for (int it = 0; it < iters; it++) { x = mm<V>(x, y); y = mm<V>(y, x); }
#pragma unroll
  for (int k = 0; k < 8; k++) { out[16 * i + k] = x.w[k]; out[16 * i + 8 + k] = y.w[k]; }
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  size_t nthr = (size_t)sms * 8 * 256;
  uint32_t *in, *o0, *o1; cudaMalloc(&in, nthr * 64); cudaMalloc(&o0, nthr * 64); cudaMalloc(&o1, nthr * 64);
  uint32_t* h = (uint32_t*)malloc(nthr * 64);
  uint64_t s = 88172645463325252ull;
  for (size_t k = 0; k < nthr * 16; k++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[k] = (uint32_t)(s >> 11); }
  cudaMemcpy(in, h, nthr * 64, cudaMemcpyHostToDevice);
  chain<0><<<sms * 8, 256>>>(in, o0, 7); chain<1><<<sms * 8, 256>>>(in, o1, 7);
  uint32_t *h0 = (uint32_t*)malloc(nthr * 64), *h1 = (uint32_t*)malloc(nthr * 64);
  cudaMemcpy(h0, o0, nthr * 64, cudaMemcpyDeviceToHost); cudaMemcpy(h1, o1, nthr * 64, cudaMemcpyDeviceToHost);
  size_t bad = 0; for (size_t k = 0; k < nthr * 16; k++) bad += h0[k] != h1[k];
  printf("variant B vs A mismatching words: %zu of %zu  (err=%s)\n", bad, nthr * 16, cudaGetErrorString(cudaGetLastError()));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int v = 0; v < 2; v++) {
    float best = 1e9;
    for (int r = 0; r < 4; r++) {
      cudaEventRecord(e0);
      if (v == 0) chain<0><<<sms * 8, 256>>>(in, o0, CHAIN); else chain<1><<<sms * 8, 256>>>(in, o1, CHAIN);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double prods = (double)nthr * CHAIN * 2;
    printf("variant %c: %8.3f ms  %.3e mont-muls/s  %.1f cycles/warp-product/SMSP (at %d MHz)\n", 'A' + v, best, prods / (best * 1e-3),
           (best * 1e-3) * clk * 1e3 / (prods / 32 / (sms * 4)), clk / 1000);
  }
  return 0;
}
