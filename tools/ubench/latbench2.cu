// Single-warp latency: k independent Montgomery multiplications per step (does ptxas / the hardware overlap them?),
// and one cached point addition (4 + 4 independent multiplications).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../dusk_zerocaf_b200/csrc/zc_point.cuh"
using namespace zc;
template <int K>
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
    for (int j = 0; j < K; j++) x[j] = mont_mul<ModP>(x[j], y);
  }
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < K; j++)
#pragma unroll
    for (int k = 0; k < 8; k++) acc ^= x[j].w[k];
  out[threadIdx.x] = acc;
}
__global__ void __launch_bounds__(32) chainAdd(const uint32_t* in, uint32_t* out, int iters) {
  Pt p; PtCached q;
  Fe* pf = &p.X; Fe* qf = &q.YpX;
#pragma unroll
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int k = 0; k < 8; k++) { pf[j].w[k] = in[64 * threadIdx.x + 8 * j + k] & (k == 7 ? 0x0fffffffu : 0xffffffffu); qf[j].w[k] = in[64 * threadIdx.x + 32 + 8 * j + k] & (k == 7 ? 0x0fffffffu : 0xffffffffu); }
#pragma unroll 1
  for (int it = 0; it < iters; it++) p = pt_add_cached(p, q);
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int k = 0; k < 8; k++) acc ^= pf[j].w[k];
  out[threadIdx.x] = acc;
}
int main() {
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  uint32_t *in, *o; cudaMalloc(&in, 32 * 256); cudaMalloc(&o, 4096);
  cudaMemset(in, 0x5b, 32 * 256);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 1000;
  for (int v = 0; v < 5; v++) {
    float ms = 0;
    for (int r = 0; r < 3; r++) {
      cudaEventRecord(e0);
      if (v == 0) chainK<1><<<1, 32>>>(in, o, iters); else if (v == 1) chainK<2><<<1, 32>>>(in, o, iters);
      else if (v == 2) chainK<4><<<1, 32>>>(in, o, iters); else if (v == 3) chainK<6><<<1, 32>>>(in, o, iters); else chainAdd<<<1, 32>>>(in, o, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
    }
    const char* nm[] = {"1 mult/step", "2 independent mults/step", "4 independent mults/step", "6 independent mults/step", "pt_add_cached (8 mults)"};
    printf("%-28s %.0f ns per step (%.0f cycles)\n", nm[v], ms * 1e6 / iters, ms * 1e-3 * clk * 1e3 / iters);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
