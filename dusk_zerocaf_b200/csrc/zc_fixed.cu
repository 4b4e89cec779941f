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
#include <stdlib.h>

#include "zc_point.cuh"
#include "zc_quad.cuh"

using namespace zc;

namespace {

constexpr int NDIG = 63;
constexpr int ENTRY_WORDS = 24;                       // (y+x, y-x, 2dxy), 8 words each
constexpr size_t TABLE_BYTES = (size_t)NDIG * 8 * ENTRY_WORDS * 4;
// In shared memory an entry occupies 28 words: the eight entries a digit can select then start in eight different
// 4-bank groups (28 k mod 32 = 0, 28, 24, ... 4), so a warp's 16-byte reads of eight different entries do not collide
// (at the natural 24-word stride entries k and k + 4 share their banks).
constexpr int ENTRY_SMEM_WORDS = 28;
constexpr size_t TABLE_SMEM_BYTES = (size_t)NDIG * 8 * ENTRY_SMEM_WORDS * 4;

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
    r = mont_sqr<ModP>(r);
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

// ---- the multiplication kernel ---------------------------------------------------------------------------------------
// The 48 KiB table is the same for every thread: each CTA stages it into shared memory with ONE TMA bulk copy
// (cp.async.bulk global -> shared, completion on an mbarrier; SASS UBLKCP) issued by an elected thread, and the grid is
// persistent (a few CTAs per SM, each looping over its scalars) so the copy is paid once per CTA, not once per 128
// scalars.  A digit then picks its entry with per-lane LDS.128 reads (32 different entries per warp: a handful of bank
// conflicts instead of 32 L1 wavefronts per load through the global path).  Measured slower than the L1 path (see the
// launch site): opt-in with ZC_FIXED_TMA, the default stays basepoint_mul_ldg_kernel.
__device__ __forceinline__ Fe ldg_fe(const uint32_t* __restrict__ p) {
  const uint4 lo = __ldg(reinterpret_cast<const uint4*>(p)), hi = __ldg(reinterpret_cast<const uint4*>(p + 4));
  return Fe{{lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w}};
}
__device__ __forceinline__ Fe lds_fe(const uint32_t* p) {
  const uint4 lo = *reinterpret_cast<const uint4*>(p), hi = *reinterpret_cast<const uint4*>(p + 4);
  return Fe{{lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w}};
}

// signed nibbles of a scalar, least significant first (same recoding as the variable-base window kernel)
__device__ __forceinline__ void signed_nibbles(const Fe& s, uint32_t (&dig)[8]) {
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
__device__ __forceinline__ int nibble_at(const uint32_t (&dig)[8], int j) {
  uint32_t nib = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) if ((j >> 3) == k) nib = dig[k];
  nib = (nib >> (4 * (j & 7))) & 15u;
  return (nib >= 8u) ? (int)nib - 16 : (int)nib;
}

constexpr int FX_TPB = 128;
__global__ void __launch_bounds__(FX_TPB) basepoint_mul_ldg_kernel(const uint64_t* __restrict__ scalars, const uint32_t* __restrict__ table,
                                                                   uint64_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * FX_TPB + threadIdx.x;
  const bool live = i < n;
  uint32_t dig[8];
  signed_nibbles(fe_load52(scalars + 5 * (live ? i : 0)), dig);
  Pt Q = pt_identity_mont();
#pragma unroll 1
  for (int j = 0; j < NDIG; j++) {
    const int d = nibble_at(dig, j);
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

// Q + (+-entry), entry = (y+x, y-x, 2dxy) with Z = 1 in shared memory; lazy linear combinations as in msm_accum_kernel
// (factors < 2m, 2m, 3m, 3m; every product < 9 m^2 < R m), canonical outputs.
__device__ __forceinline__ Pt fx_add_smem(const Pt& p, const uint32_t* e, bool neg) {
  const Fe zero{{0, 0, 0, 0, 0, 0, 0, 0}};
  Fe A = fx_mul(fe_sub_lazy<1>(p.Y, p.X), lds_fe(e + (neg ? 0 : 8)));
  Fe B = fx_mul(fe_add_lazy(p.Y, p.X), lds_fe(e + (neg ? 8 : 0)));
  Fe t2d = lds_fe(e + 16);
  if (neg) t2d = fe_sub_lazy<1>(zero, t2d);
  Fe C = fx_mul(p.T, t2d);
  Fe D = fe_dbl_lazy(p.Z);
  Fe E = fe_sub_lazy<1>(B, A), F = fe_sub_lazy<1>(D, C), G = fe_add_lazy(D, C), H = fe_add_lazy(B, A);
  Pt r;
  r.X = fx_mul(E, F); r.Y = fx_mul(G, H); r.Z = fx_mul(F, G); r.T = fx_mul(E, H);
  return r;
}

// BT threads per CTA, MINB CTAs per SM; INL: the round-1 inlined addition (pt_add_affcached, 96 registers) on the staged
// entry instead of the out-of-line multiplier + lazy combinations.
template <int BT, int MINB, bool INL>
__global__ void __launch_bounds__(BT, MINB) basepoint_mul_kernel(const uint64_t* __restrict__ scalars, const uint32_t* __restrict__ table,
                                                               uint64_t* __restrict__ out, size_t n) {
  extern __shared__ __align__(128) uint32_t fx_table[];             // TABLE_SMEM_BYTES
  __shared__ __align__(8) uint64_t fx_bar;
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&fx_bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)TABLE_BYTES) : "memory");
  __syncthreads();
  // one 96-byte bulk copy per entry (504 of them, spread over the first warps), all completing on the same barrier
  for (int e = threadIdx.x; e < NDIG * 8; e += BT)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(fx_table + (size_t)e * ENTRY_SMEM_WORDS)), "l"(table + (size_t)e * ENTRY_WORDS),
                   "r"((uint32_t)(ENTRY_WORDS * 4)), "r"(bar) : "memory");
  // the first scalar's digits are recoded while the table is in flight
  size_t i = (size_t)blockIdx.x * BT + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * BT;
  uint32_t dig[8];
  signed_nibbles(fe_load52(scalars + 5 * (i < n ? i : 0)), dig);
  {
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar) : "memory");
  }
  // whole warps stay in the loop together (the addition is warp-uniform code with __any_sync)
  const size_t i_warp0 = i - (threadIdx.x & 31);
#pragma unroll 1
  for (size_t base = i_warp0; base < n; base += stride, i += stride) {
    const bool live = i < n;
    if (base != i_warp0) signed_nibbles(fe_load52(scalars + 5 * (live ? i : 0)), dig);
    Pt Q = pt_identity_mont();
#pragma unroll 1
    for (int j = 0; j < NDIG; j++) {
      const int d = nibble_at(dig, j);
      const int mag = d < 0 ? -d : d;
      if (__any_sync(0xffffffffu, mag != 0)) {
        const uint32_t* e = fx_table + (size_t)(j * 8 + (mag ? mag - 1 : 0)) * ENTRY_SMEM_WORDS;
        Pt r;
        if (INL) { PtAffCached c{lds_fe(e), lds_fe(e + 8), lds_fe(e + 16)}; r = pt_add_affcached(Q, c, d < 0); }
        else r = fx_add_smem(Q, e, d < 0);
        if (mag != 0) Q = r;
      }
    }
    if (live) pt_store52(out + 20 * i, pt_from_mont(Q));
  }
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
  // Default: the table is read through L1 (__ldg).  ZC_FIXED_TMA = 0 / 1 / 2 selects the TMA-staged shared-memory variants
  // (192 x 3 with the out-of-line multiplier, 320 x 2 and 192 x 3 with the inlined addition).  Measured on B200 at 2^20
  // scalars (profiles/r02_fixed_base_tma_ab.txt): 6.82 ms through L1; 7.49 / 7.14 / 7.20 ms with ONE 48 KiB bulk copy per
  // CTA (128 x 4 / 320 x 2 / 128 x 4 threads x CTAs); 8.04 / 7.15 / 7.78 ms with per-entry copies into the conflict-free
  // 28-word layout below.  The kernel is multiplier-bound and the L1 path keeps 20 warps per SM in 96 registers with no
  // table footprint in shared memory, so staging buys nothing here; the variants stay for A/B runs.
  static const int tma_variant = getenv("ZC_FIXED_TMA") ? atoi(getenv("ZC_FIXED_TMA")) : -1;
  const size_t nblk = (n + FX_TPB - 1) / FX_TPB;
  if (tma_variant < 0) {
    basepoint_mul_ldg_kernel<<<(unsigned)nblk, FX_TPB, 0, ctx->stream>>>(scalars, (const uint32_t*)ctx->basepoint_table, out, n);
  } else {
#define ZC_FX_LAUNCH(BT, MINB, INL) do {                                                                                        \
      static bool attr_set = false;                                                                                             \
      if (!attr_set) { ZC_CUDA(ctx, cudaFuncSetAttribute(basepoint_mul_kernel<BT, MINB, INL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TABLE_SMEM_BYTES)); attr_set = true; } \
      const size_t nb_ = (n + BT - 1) / BT, cap_ = (size_t)MINB * ctx->sm_count;                                                 \
      basepoint_mul_kernel<BT, MINB, INL><<<(unsigned)(nb_ < cap_ ? nb_ : cap_), BT, TABLE_SMEM_BYTES, ctx->stream>>>(scalars, (const uint32_t*)ctx->basepoint_table, out, n); \
    } while (0)
    if (tma_variant == 1) ZC_FX_LAUNCH(320, 2, true);
    else if (tma_variant == 2) ZC_FX_LAUNCH(192, 3, true);
    else ZC_FX_LAUNCH(192, 3, false);
#undef ZC_FX_LAUNCH
  }
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
