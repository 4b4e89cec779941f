// zc_fixed.cu -- fixed-base scalar multiplication  out[i] = [s_i] B  for the curve's basepoint.
//
// Reference: `&BASEPOINT * &scalar` -- Mul<&Scalar> for &EdwardsPoint, /root/reference/src/edwards.rs:547-577 (LSB-first
// double_and_add), BASEPOINT /root/reference/src/backend/u64/constants.rs:188-211.  The reference's own fixed-base
// attempt, window_naf_mul (edwards.rs:155-171) over BASEPOINT_ODD_MULTIPLES_TABLE (constants.rs:216-972), indexes its
// table wrongly and is untested; this is the working equivalent: the same group element through a
// table built on the device, compared canonically (affine / Ristretto encoding), never limb-wise.
//
// Signed radix-16 digits  s = sum_j d_j 16^j,  d_j in [-8, 8),  63 digits (s < L < 2^250).  The table holds
// (k+1) 16^j B  for j < 63, k < 8 in AFFINE cached form (y+x, y-x, 2dxy) -- 48 KiB, built once per context -- so one
// scalar multiplication is 63 seven-multiplication additions and no doubling: ~440 field multiplications instead of
// ~2600 for the variable-base window method.
#include "zc_internal.h"
#include "zc_point.cuh"

using namespace zc;

namespace {

constexpr int NDIG = 63;
constexpr int ENTRY_WORDS = 24;                       // (y+x, y-x, 2dxy), 8 words each
constexpr size_t TABLE_BYTES = (size_t)NDIG * 8 * ENTRY_WORDS * 4;

__device__ __forceinline__ Pt BASEPOINT_MONT() {     // constants.rs:188-211, coordinates times R
  Pt b;
  b.X = Fe{{0x1ae76da9u, 0x038daf26u, 0xe2edf90au, 0xb6417543u, 0xae728baau, 0x75285a28u, 0xb8c8f151u, 0x0d0a23acu}};
  b.Y = Fe{{0x962c6b19u, 0x2a865c6eu, 0x041ba421u, 0x6f0339a0u, 0x33333332u, 0x33333333u, 0x33333333u, 0x03333333u}};
  b.Z = Consts<ModP>::R1();
  b.T = Fe{{0x102474ffu, 0xceee9c4au, 0x21c1fbd2u, 0x6d5a798fu, 0x3577ed66u, 0xdfe502e5u, 0x6edef730u, 0x07d2e234u}};
  return b;
}

__device__ __noinline__ Fe fx_mul(Fe a, Fe b) { return mont_mul<ModP>(a, b); }

// a^(p-2), Montgomery form (one-time table construction only)
__device__ Fe fx_invert(const Fe& a) {
  const uint32_t e[8] = {0x5cf5d3ebu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0u, 0u, 0u, 0x10000000u};
  Fe r = a;
#pragma unroll 1
  for (int bit = 251; bit >= 0; bit--) {
    r = fx_mul(r, r);
    if ((e[bit >> 5] >> (bit & 31)) & 1u) r = fx_mul(r, a);
  }
  return r;
}

// thread j builds the eight multiples of 16^j B
__global__ void __launch_bounds__(64) basepoint_table_kernel(uint32_t* __restrict__ table) {
  typedef ModP M;
  const int j = threadIdx.x;
  if (j >= NDIG) return;
  Pt b = BASEPOINT_MONT();
#pragma unroll 1
  for (int i = 0; i < 4 * j; i++) b = pt_double_fast(b);
  const PtCached cb = pt_to_cached(b);
  Pt acc = b;
#pragma unroll 1
  for (int k = 0; k < 8; k++) {
    if (k > 0) acc = pt_add_cached(acc, cb);
    const Fe zi = fx_invert(acc.Z);
    const Fe x = fx_mul(acc.X, zi), y = fx_mul(acc.Y, zi);
    const Fe ypx = fe_add<M>(y, x), ymx = fe_sub<M>(y, x), t2d = fx_mul(fx_mul(x, y), D2_MONT());
    uint32_t* o = table + (size_t)(j * 8 + k) * ENTRY_WORDS;
#pragma unroll
    for (int w = 0; w < 8; w++) { o[w] = ypx.w[w]; o[8 + w] = ymx.w[w]; o[16 + w] = t2d.w[w]; }
  }
}

__device__ __forceinline__ Fe ldg_fe(const uint32_t* __restrict__ p) {
  const uint4 lo = __ldg(reinterpret_cast<const uint4*>(p)), hi = __ldg(reinterpret_cast<const uint4*>(p + 4));
  return Fe{{lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w}};
}

constexpr int FX_TPB = 128;
__global__ void __launch_bounds__(FX_TPB) basepoint_mul_kernel(const uint64_t* __restrict__ scalars, const uint32_t* __restrict__ table,
                                                               uint64_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * FX_TPB + threadIdx.x;
  const bool live = i < n;
  Fe s = fe_load52(scalars + 5 * (live ? i : 0));
  // signed nibbles, least significant first (same recoding as the variable-base window kernel)
  uint32_t dig[8];
  {
    uint32_t carry = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      uint32_t w = s.w[k], o = 0;
#pragma unroll
      for (int jj = 0; jj < 8; jj++) {
        uint32_t d = ((w >> (4 * jj)) & 15u) + carry;
        carry = (d >= 8u) ? 1u : 0u;
        o |= (d & 15u) << (4 * jj);
      }
      dig[k] = o;
    }
  }
  Pt Q = pt_identity_mont();
#pragma unroll 1
  for (int j = 0; j < NDIG; j++) {
    uint32_t nib = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) if ((j >> 3) == k) nib = dig[k];
    nib = (nib >> (4 * (j & 7))) & 15u;
    const int d = (nib >= 8u) ? (int)nib - 16 : (int)nib;
    const int mag = d < 0 ? -d : d;
    if (__any_sync(0xffffffffu, mag != 0)) {
      const uint32_t* e = table + (size_t)(j * 8 + (mag ? mag - 1 : 0)) * ENTRY_WORDS;
      PtAffCached c{ldg_fe(e), ldg_fe(e + 8), ldg_fe(e + 16)};
      Pt r = pt_add_affcached(Q, c, d < 0);
      if (mag != 0) Q = r;
    }
  }
  if (live) pt_store52(out + 20 * i, pt_from_mont(Q));
}

}  // namespace

extern "C" {

int32_t zc_basepoint_mul_batch_dev(zc_ctx* ctx, const uint64_t* scalars, uint64_t* out, size_t n) {
  if (!ctx) return ZC_ERR_NULL;
  ZC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n > ((size_t)1 << 31)) return zc_fail(ctx, ZC_ERR_SIZE, "n exceeds 2^31");
  if (n == 0) return ZC_OK;
  if (!scalars || !out) return zc_fail(ctx, ZC_ERR_NULL, "null pointer argument");
  if (!ctx->basepoint_table) {
    ZC_CUDA(ctx, cudaMalloc(&ctx->basepoint_table, TABLE_BYTES));
    basepoint_table_kernel<<<1, 64, 0, ctx->stream>>>((uint32_t*)ctx->basepoint_table);
    ctx->launches++;
  }
  basepoint_mul_kernel<<<(unsigned)((n + FX_TPB - 1) / FX_TPB), FX_TPB, 0, ctx->stream>>>(scalars, (const uint32_t*)ctx->basepoint_table, out, n);
  ctx->launches++;
  ZC_CUDA(ctx, cudaGetLastError());
  return ZC_OK;
}

int32_t zc_basepoint_mul_batch(zc_ctx* ctx, const uint64_t* scalars, uint64_t* out, size_t n) {
  if (!ctx) return ZC_ERR_NULL;
  ZC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n > ((size_t)1 << 31)) return zc_fail(ctx, ZC_ERR_SIZE, "n exceeds 2^31");
  if (n == 0) return ZC_OK;
  if (!scalars || !out) return zc_fail(ctx, ZC_ERR_NULL, "null pointer argument");
  void *ds = nullptr, *dout = nullptr;
  int32_t rc;
  if ((rc = zc_scratch(ctx, 1, n * 40, &ds))) return rc;
  if ((rc = zc_scratch(ctx, 2, n * 160, &dout))) return rc;
  ZC_CUDA(ctx, cudaMemcpyAsync(ds, scalars, n * 40, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = zc_basepoint_mul_batch_dev(ctx, (const uint64_t*)ds, (uint64_t*)dout, n))) return rc;
  ZC_CUDA(ctx, cudaMemcpyAsync(out, dout, n * 160, cudaMemcpyDeviceToHost, ctx->stream));
  ZC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZC_OK;
}

}  // extern "C"
