// karabench.cu -- round-2 experiment, NOT part of the product: does a one-level Karatsuba split of the 8 x 8-word product
// (3 x 4x4 = 48 wide multiplies instead of 64) pay in the config-2 kernel, which is bound by the IMAD.WIDE pipe (196 wide
// multiplies per pair, pipe 79 % busy -- profiles/r01_fe_mul_square_ncu.csv)?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o karabench karabench.cu
//   ./karabench            # 1. checks mul_wide_8x8_kara against mul_wide_8x8 on 2^20 random operand pairs
//                          # 2. times the config-2 kernel body (2^24 pairs) with the schoolbook and the Karatsuba product
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../../dusk_zerocaf_b200/csrc/zc_fe.cuh"

using namespace zc;

// t[0..7] = a[0..3] * b[0..3]: 16 wide multiplies, row by row on 64-bit accumulators (ptxas emits IMAD.WIDE + IADD3.X)
__device__ __forceinline__ void mul_wide_4x4(uint32_t (&t)[8], const uint32_t* a, const uint32_t* b) {
#pragma unroll
  for (int k = 0; k < 8; k++) t[k] = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint64_t carry = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint64_t p = (uint64_t)a[j] * b[i] + t[i + j] + carry;
      t[i + j] = (uint32_t)p;
      carry = p >> 32;
    }
    t[i + 4] = (uint32_t)carry;
  }
}

// t[0..15] = a * b with one Karatsuba level:  z0 = a0 b0, z2 = a1 b1, z1 = (a0 + a1)(b0 + b1) - z0 - z2
__device__ __forceinline__ void mul_wide_8x8_kara(uint32_t (&t)[16], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
  uint32_t z0[8], z2[8], zm[8], sa[4], sb[4];
  mul_wide_4x4(z0, a, b);
  mul_wide_4x4(z2, a + 4, b + 4);
  uint64_t c = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) { c += (uint64_t)a[k] + a[4 + k]; sa[k] = (uint32_t)c; c >>= 32; }
  const uint32_t ca = (uint32_t)c;
  c = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) { c += (uint64_t)b[k] + b[4 + k]; sb[k] = (uint32_t)c; c >>= 32; }
  const uint32_t cb = (uint32_t)c;
  mul_wide_4x4(zm, sa, sb);
  // z1 = zm + (ca ? sb : 0) 2^128 + (cb ? sa : 0) 2^128 + (ca & cb) 2^256 - z0 - z2      (9 words, non-negative)
  uint32_t z1[9];
  const uint32_t ma = 0u - ca, mb = 0u - cb;
  c = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) z1[k] = zm[k];
#pragma unroll
  for (int k = 0; k < 4; k++) { c += (uint64_t)zm[4 + k] + (sb[k] & ma) + (sa[k] & mb); z1[4 + k] = (uint32_t)c; c >>= 32; }
  z1[8] = (uint32_t)c + (ca & cb);
  int64_t bw = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    bw += (int64_t)z1[k] - z0[k] - z2[k];
    z1[k] = (uint32_t)bw;
    bw >>= 32;                                  // arithmetic shift: borrow of up to -2
  }
  z1[8] = (uint32_t)((int64_t)z1[8] + bw);
  // t = z0 + z1 2^128 + z2 2^256
#pragma unroll
  for (int k = 0; k < 4; k++) t[k] = z0[k];
  c = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) { c += (uint64_t)z0[4 + k] + z1[k]; t[4 + k] = (uint32_t)c; c >>= 32; }
#pragma unroll
  for (int k = 0; k < 4; k++) { c += (uint64_t)z2[k] + z1[4 + k]; t[8 + k] = (uint32_t)c; c >>= 32; }
  c += z1[8];
#pragma unroll
  for (int k = 0; k < 4; k++) { c += (uint64_t)z2[4 + k]; t[12 + k] = (uint32_t)c; c >>= 32; }
}

__device__ __forceinline__ uint32_t xs(uint64_t& s) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); }

__global__ void check_kernel(unsigned long long* bad, int rounds) {
  uint64_t s = 0x9e3779b97f4a7c15ull * (blockIdx.x * blockDim.x + threadIdx.x + 1);
  for (int r = 0; r < rounds; r++) {
    uint32_t a[8], b[8], t0[16], t1[16];
    for (int k = 0; k < 8; k++) { a[k] = xs(s); b[k] = xs(s); }
    if (r == 0) for (int k = 0; k < 8; k++) { a[k] = 0xffffffffu; b[k] = 0xffffffffu; }     // carries everywhere
    if (r == 1) for (int k = 0; k < 8; k++) { a[k] = k < 4 ? 0xffffffffu : 1u; b[k] = k < 4 ? 1u : 0xffffffffu; }
    mul_wide_8x8(t0, a, b);
    mul_wide_8x8_kara(t1, a, b);
    bool eq = true;
    for (int k = 0; k < 16; k++) eq = eq && (t0[k] == t1[k]);
    if (!eq) atomicAdd(bad, 1ull);
  }
}

template <int VARIANT>
__global__ void __launch_bounds__(256) mul_square_kernel(const uint64_t* __restrict__ a, const uint64_t* __restrict__ b,
                                                         uint64_t* __restrict__ prod, uint64_t* __restrict__ sq, size_t n) {
  typedef ModP M;
  size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const Fe x = fe_load52_shl<Shape<M>::SA>(a + 5 * i);
  const Fe y = fe_load52_shl<Shape<M>::SB>(b + 5 * i);
  uint32_t t[16];
  if (VARIANT == 0) mul_wide_8x8(t, x.w, y.w); else mul_wide_8x8_kara(t, x.w, y.w);
  fe_store52(prod + 5 * i, fold_product<M, Shape<M>::S>(t));
  fe_store52(sq + 5 * i, fe_sqr_normal_pre<M>(x));
}

int main() {
  unsigned long long* bad;
  cudaMallocManaged(&bad, 8);
  *bad = 0;
  check_kernel<<<1024, 256>>>(bad, 4);
  cudaDeviceSynchronize();
  printf("karatsuba vs schoolbook on 2^20 operand pairs: %llu mismatches\n", *bad);
  const size_t n = (size_t)1 << 24;
  uint64_t *a, *b, *p, *s;
  cudaMalloc(&a, n * 40); cudaMalloc(&b, n * 40); cudaMalloc(&p, n * 40); cudaMalloc(&s, n * 40);
  cudaMemset(a, 0x5a, n * 40); cudaMemset(b, 0x3c, n * 40);          // limbs above 2^52: fine for timing, not canonical
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int variant = 0; variant < 2; variant++) {
    float best = 1e9f;
    for (int rep = 0; rep < 8; rep++) {
      cudaEventRecord(e0);
      if (variant == 0) mul_square_kernel<0><<<(unsigned)(n / 256), 256>>>(a, b, p, s, n);
      else mul_square_kernel<1><<<(unsigned)(n / 256), 256>>>(a, b, p, s, n);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep >= 2 && ms < best) best = ms;
    }
    printf("config-2 body, %s product: %.3f ms per 2^24 pairs\n", variant ? "Karatsuba (48 wide multiplies)" : "schoolbook (64 wide multiplies)", best);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return *bad != 0;
}
