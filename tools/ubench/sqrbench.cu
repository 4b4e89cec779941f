// Dedicated Montgomery squaring (mont_sqr: 36 + 32 wide multiplies) against mont_mul(a, a) (64 + 32): bit-exact check on
// random and edge inputs for both moduli, then throughput of dependent chains at full occupancy.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../dusk_zerocaf_b200/csrc/zc_fe.cuh"
using namespace zc;
#define CHAIN 512

template <class M, int V> __device__ __forceinline__ Fe sq(const Fe& a) {
  if (V == 0) return mont_mul<M>(a, a);
  else return mont_sqr<M>(a);
}
template <class M>
__global__ void check(const uint32_t* in, uint32_t* bad, size_t n, uint32_t topmask) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fe x;
#pragma unroll
  for (int k = 0; k < 8; k++) x.w[k] = in[8 * i + k];
  x.w[7] &= topmask;
  Fe a = mont_mul<M>(x, x), b = mont_sqr<M>(x);
  Fe al = mont_mul_lazy<M>(x, x), bl = mont_sqr_lazy<M>(x);
  uint32_t d = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) d |= (a.w[k] ^ b.w[k]) | (al.w[k] ^ bl.w[k]);
  if (d) atomicAdd(bad, 1u);
}
template <class M, int V>
__global__ void __launch_bounds__(256) chain(const uint32_t* in, uint32_t* out, int iters) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  Fe x;
#pragma unroll
  for (int k = 0; k < 8; k++) x.w[k] = in[8 * i + k];
  x.w[7] &= 0x0fffffffu;
#pragma unroll 1
  for (int it = 0; it < iters; it++) x = sq<M, V>(x);
#pragma unroll
  for (int k = 0; k < 8; k++) out[8 * i + k] = x.w[k];
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  size_t nthr = (size_t)sms * 8 * 256;
  uint32_t *in, *o0, *o1, *bad; cudaMalloc(&in, nthr * 32); cudaMalloc(&o0, nthr * 32); cudaMalloc(&o1, nthr * 32); cudaMalloc(&bad, 4);
  uint32_t* h = (uint32_t*)malloc(nthr * 32);
  uint64_t s = 88172645463325252ull;
  for (size_t k = 0; k < nthr * 8; k++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[k] = (uint32_t)(s >> 11); }
  // edge values in the first rows: 0, 1, 2^32 - 1 words, p - 1, p, 2p - 1, L - 1, L, all ones
  const uint32_t edge[][8] = {
    {0, 0, 0, 0, 0, 0, 0, 0}, {1, 0, 0, 0, 0, 0, 0, 0}, {0xffffffffu, 0, 0, 0, 0, 0, 0, 0},
    {0x5cf5d3ecu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0, 0, 0, 0x10000000u},
    {0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0, 0, 0, 0x10000000u},
    {0xb9eba7d9u, 0xb024c634u, 0x45ef39acu, 0x29bdf3bdu, 0, 0, 0, 0x20000000u},
    {0x755fc862u, 0x6ab4036fu, 0x822fd593u, 0x0ae6c74du, 0, 0, 0, 0x02000000u},
    {0x755fc863u, 0x6ab4036fu, 0x822fd593u, 0x0ae6c74du, 0, 0, 0, 0x02000000u},
    {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu},
    {0, 0, 0, 0, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu}, {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0, 0, 0, 0}};
  for (size_t e = 0; e < sizeof(edge) / 32; e++) for (int k = 0; k < 8; k++) h[8 * e + k] = edge[e][k];
  cudaMemcpy(in, h, nthr * 32, cudaMemcpyHostToDevice);
  // the squaring must agree with the product for every input the callers can feed: canonical, lazy (< 4m < 2^255), and here
  // even full 256-bit values for mod p (mask 0xffffffff) since both compute (a^2 + Q m) / R with the same Q
  struct { const char* name; uint32_t mask; } cases[] = {{"< 2^252", 0x0fffffffu}, {"< 2^254", 0x3fffffffu}, {"< 2^255", 0x7fffffffu}, {"< 2^256", 0xffffffffu}};
  for (auto& c : cases) {
    uint32_t hb = 0;
    cudaMemset(bad, 0, 4);
    check<ModP><<<(unsigned)((nthr + 255) / 256), 256>>>(in, bad, nthr, c.mask);
    cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
    printf("mod p, inputs %s: %u of %zu mismatch (canonical and lazy results)\n", c.name, hb, nthr);
    cudaMemset(bad, 0, 4);
    check<ModL><<<(unsigned)((nthr + 255) / 256), 256>>>(in, bad, nthr, c.mask);
    cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
    printf("mod L, inputs %s: %u of %zu mismatch\n", c.name, hb, nthr);
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int v = 0; v < 2; v++) {
    float best = 1e9;
    for (int r = 0; r < 4; r++) {
      cudaEventRecord(e0);
      if (v == 0) chain<ModP, 0><<<sms * 8, 256>>>(in, o0, CHAIN); else chain<ModP, 1><<<sms * 8, 256>>>(in, o1, CHAIN);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double prods = (double)nthr * CHAIN;
    printf("%s: %8.3f ms  %.3e squarings/s  %.1f cycles/warp-squaring/SMSP (at %d MHz)\n", v ? "mont_sqr      " : "mont_mul(a, a)", best, prods / (best * 1e-3),
           (best * 1e-3) * clk * 1e3 / (prods / 32 / (sms * 4)), clk / 1000);
  }
  // one warp alone: latency of a dependent squaring
  for (int v = 0; v < 2; v++) {
    cudaEventRecord(e0);
    if (v == 0) chain<ModP, 0><<<1, 32>>>(in, o0, 4096); else chain<ModP, 1><<<1, 32>>>(in, o1, 4096);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%s: one warp, dependent chain: %.1f ns per squaring\n", v ? "mont_sqr      " : "mont_mul(a, a)", ms * 1e6 / 4096);
  }
  uint32_t *h0 = (uint32_t*)malloc(nthr * 32), *h1 = (uint32_t*)malloc(nthr * 32);
  chain<ModP, 0><<<sms * 8, 256>>>(in, o0, 33); chain<ModP, 1><<<sms * 8, 256>>>(in, o1, 33);
  cudaMemcpy(h0, o0, nthr * 32, cudaMemcpyDeviceToHost); cudaMemcpy(h1, o1, nthr * 32, cudaMemcpyDeviceToHost);
  size_t badw = 0; for (size_t k = 0; k < nthr * 8; k++) badw += h0[k] != h1[k];
  printf("33-step chains: %zu mismatching words (err=%s)\n", badw, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
