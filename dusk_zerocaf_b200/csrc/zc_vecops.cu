// zc_vecops.cu -- the caller side of the MSM in an inner-product argument: scalar- and field-vector
// operations beyond + - x, the signed-digit recodings, and the 32-byte wire format of both residue types.
//
//   zc_{fe,scalar}_pow_batch         Pow            /root/reference/src/backend/u64/field.rs:334-354, scalar.rs:293-322
//   zc_{fe,scalar}_half_batch        Half           field.rs:317-323, scalar.rs:285-291   (a * 2^-1 mod m)
//   zc_{fe,scalar}_to_bytes_batch    to_bytes       field.rs:591-631, scalar.rs:477-516
//   zc_fe_from_bytes_batch           from_bytes     field.rs:563-587  (all 256 bits kept, no reduction)
//   zc_scalar_from_bytes_batch       from_bytes     scalar.rs:445-467 (the reference asserts <= L - 1; here ok[i] = 0)
//   zc_scalar_window_naf_batch       compute_window_NAF scalar.rs:396-415 (width 2 = compute_NAF :370-390)
//   zc_fe_sqrt_ratio_i_batch         sqrt_ratio_i   field.rs:443-491  -- lives in zc_encode.cu with the square-root chain
//
// The reference drives pow by halving the exponent and the recodings by repeated subtraction on 5-limb values; each
// returns a uniquely defined value, computed here on 8 x u32 words with the loops every lane of a warp shares.
#include "zc_internal.h"
#include "zc_fe.cuh"

using namespace zc;

namespace {

constexpr int TPB = 128;
constexpr size_t MAX_N = (size_t)1 << 31;
inline unsigned grid_for(size_t n) { return (unsigned)((n + TPB - 1) / TPB); }

template <class M>
__device__ __forceinline__ Fe modulus() { return Fe{{M::M0, M::M1, M::M2, M::M3, 0u, 0u, 0u, M::M7}}; }

template <class M>
__device__ __noinline__ Fe vmul(Fe a, Fe b) { return mont_mul<M>(a, b); }
template <class M>
__device__ __noinline__ Fe vsqr(Fe a) { return mont_sqr<M>(a); }

// a^e mod m, left-to-right over the 253 exponent bits (0^0 = 1 as in the reference: the loop never runs)
template <class M>
__global__ void __launch_bounds__(TPB) pow_kernel(const uint64_t* __restrict__ a, const uint64_t* __restrict__ e,
                                                  uint64_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  const Fe x = to_mont<M>(fe_load52(a + 5 * i));
  const Fe ex = fe_load52(e + 5 * i);
  Fe r = Consts<M>::R1();
#pragma unroll 1
  for (int bit = 252; bit >= 0; bit--) {
    r = vsqr<M>(r);
    uint32_t word = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) if ((bit >> 5) == k) word = ex.w[k];
    const Fe t = vmul<M>(r, x);
    if ((word >> (bit & 31)) & 1u) r = t;
  }
  fe_store52(out + 5 * i, from_mont<M>(r));
}

// (a + (a odd ? m : 0)) >> 1  =  a * 2^-1 mod m
template <class M>
__global__ void __launch_bounds__(TPB) half_kernel(const uint64_t* __restrict__ a, uint64_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  Fe x = fe_load52(a + 5 * i);
  const Fe m = modulus<M>();
  const uint32_t odd = 0u - (x.w[0] & 1u);
  uint64_t carry = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const uint64_t t = (uint64_t)x.w[k] + (m.w[k] & odd) + carry;
    x.w[k] = (uint32_t)t;
    carry = t >> 32;
  }
#pragma unroll
  for (int k = 0; k < 7; k++) x.w[k] = (x.w[k] >> 1) | (x.w[k + 1] << 31);
  x.w[7] >>= 1;
  fe_store52(out + 5 * i, x);
}

// the 32-byte encoding is the little-endian value itself (both byte schedules are plain bit packings)
__global__ void __launch_bounds__(TPB) to_bytes_kernel(const uint64_t* __restrict__ a, uint8_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  const Fe x = fe_load52(a + 5 * i);
  uint32_t* o = reinterpret_cast<uint32_t*>(out + 32 * i);
#pragma unroll
  for (int k = 0; k < 8; k++) o[k] = x.w[k];
}
// CHECK_L: ok[i] = (value <= L - 1), the condition Scalar::from_bytes asserts
template <bool CHECK_L>
__global__ void __launch_bounds__(TPB) from_bytes_kernel(const uint8_t* __restrict__ in, uint64_t* __restrict__ out,
                                                         uint8_t* __restrict__ ok, size_t n) {
  size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  const uint32_t* w = reinterpret_cast<const uint32_t*>(in + 32 * i);
  Fe x{{w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]}};
  fe_store52(out + 5 * i, x);
  if (CHECK_L) {
    const Fe m = modulus<ModL>();
    uint64_t borrow = 0;                      // x - L borrows  <=>  x < L
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const uint64_t t = (uint64_t)x.w[k] - m.w[k] - borrow;
      borrow = (t >> 32) & 1u;
    }
    ok[i] = (uint8_t)borrow;
  }
}

// width-w non-adjacent form, 256 signed digits per scalar, least significant first:
//   while k >= 1:  if k odd: d = k mods 2^w, k -= d (mod L);  k >>= 1
__global__ void __launch_bounds__(TPB) window_naf_kernel(const uint64_t* __restrict__ a, int width, int8_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  Fe k = fe_load52(a + 5 * i);
  const uint32_t wmask = (1u << width) - 1u, half = 1u << (width - 1);
  uint32_t* o = reinterpret_cast<uint32_t*>(out + 256 * i);
#pragma unroll 1
  for (int j = 0; j < 64; j++) {
    uint32_t packed = 0;
#pragma unroll 1
    for (int b = 0; b < 4; b++) {
      int32_t d = 0;
      if (k.w[0] & 1u) {
        const uint32_t r = k.w[0] & wmask;
        d = r >= half ? (int32_t)r - (int32_t)(wmask + 1u) : (int32_t)r;
        // k -= d:  d > 0 clears the low bits (no borrow: k >= r);  d < 0 adds |d| and may reach L (Scalar arithmetic wraps)
        if (d > 0) {
          k.w[0] -= (uint32_t)d;
        } else {
          uint64_t carry = (uint64_t)(-d);
#pragma unroll
          for (int q = 0; q < 8; q++) { const uint64_t t = (uint64_t)k.w[q] + carry; k.w[q] = (uint32_t)t; carry = t >> 32; }
          reduce_once<ModL>(k);
        }
      }
      packed |= ((uint32_t)d & 0xffu) << (8 * b);
#pragma unroll
      for (int q = 0; q < 7; q++) k.w[q] = (k.w[q] >> 1) | (k.w[q + 1] << 31);
      k.w[7] >>= 1;
    }
    o[j] = packed;
  }
}

// Scalar::into_bits (scalar.rs:352-366): the 256 bits of the value, least significant first, one byte each.
// One thread per (scalar, 16-bit group): 16-byte stores, coalesced.
__global__ void __launch_bounds__(TPB) into_bits_kernel(const uint64_t* __restrict__ a, uint8_t* __restrict__ out, size_t n) {
  size_t g = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (g >= 16 * n) return;
  const size_t i = g >> 4;
  const int grp = (int)(g & 15);
  const Fe x = fe_load52(a + 5 * i);
  const uint32_t bits = (x.w[grp >> 1] >> (16 * (grp & 1))) & 0xffffu;
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint32_t nib = (bits >> (4 * k)) & 15u;
    o[k] = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
  }
  *reinterpret_cast<uint4*>(out + 256 * i + 16 * grp) = make_uint4(o[0], o[1], o[2], o[3]);
}

// small synchronous host wrappers
template <class F>
int32_t host_run(zc_ctx* ctx, const void* in0, size_t in0_bytes, const void* in1, size_t in1_bytes, void* out, size_t out_bytes,
                 void* out2, size_t out2_bytes, F run) {
  void *d0 = nullptr, *d1 = nullptr, *dout = nullptr, *dout2 = nullptr;
  int32_t rc;
  if ((rc = zc_scratch(ctx, 0, in0_bytes, &d0))) return rc;
  if (in1 && (rc = zc_scratch(ctx, 1, in1_bytes, &d1))) return rc;
  if ((rc = zc_scratch(ctx, 2, out_bytes, &dout))) return rc;
  if (out2 && (rc = zc_scratch(ctx, 3, out2_bytes, &dout2))) return rc;
  ZC_CUDA(ctx, cudaMemcpyAsync(d0, in0, in0_bytes, cudaMemcpyHostToDevice, ctx->stream));
  if (in1) ZC_CUDA(ctx, cudaMemcpyAsync(d1, in1, in1_bytes, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = run(d0, d1, dout, dout2))) return rc;
  ZC_CUDA(ctx, cudaMemcpyAsync(out, dout, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  if (out2) ZC_CUDA(ctx, cudaMemcpyAsync(out2, dout2, out2_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  ZC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZC_OK;
}

}  // namespace

#define ZC_VEC_PROLOGUE(ctx, n, cond)                                                           \
  do {                                                                                          \
    if (!(ctx)) return ZC_ERR_NULL;                                                             \
    ZC_CUDA(ctx, cudaSetDevice((ctx)->device));                                                 \
    if ((n) > MAX_N) return zc_fail(ctx, ZC_ERR_SIZE, "n exceeds 2^31");                        \
    if ((n) == 0) return ZC_OK;                                                                 \
    if (!(cond)) return zc_fail(ctx, ZC_ERR_NULL, "null pointer argument");                     \
  } while (0)
#define ZC_VEC_LAUNCHED(ctx)              \
  do {                                    \
    (ctx)->launches++;                    \
    ZC_CUDA(ctx, cudaGetLastError());     \
    return ZC_OK;                         \
  } while (0)

extern "C" {

#define ZC_DEFINE_POW_HALF(name, MOD)                                                                                     \
  int32_t zc_##name##_pow_batch_dev(zc_ctx* ctx, const uint64_t* a, const uint64_t* e, uint64_t* out, size_t n) {         \
    ZC_VEC_PROLOGUE(ctx, n, a && e && out);                                                                               \
    pow_kernel<MOD><<<grid_for(n), TPB, 0, ctx->stream>>>(a, e, out, n);                                                  \
    ZC_VEC_LAUNCHED(ctx);                                                                                                 \
  }                                                                                                                       \
  int32_t zc_##name##_pow_batch(zc_ctx* ctx, const uint64_t* a, const uint64_t* e, uint64_t* out, size_t n) {             \
    ZC_VEC_PROLOGUE(ctx, n, a && e && out);                                                                               \
    return host_run(ctx, a, n * 40, e, n * 40, out, n * 40, nullptr, 0, [&](void* d0, void* d1, void* dout, void*) {      \
      return zc_##name##_pow_batch_dev(ctx, (const uint64_t*)d0, (const uint64_t*)d1, (uint64_t*)dout, n);                \
    });                                                                                                                   \
  }                                                                                                                       \
  int32_t zc_##name##_half_batch_dev(zc_ctx* ctx, const uint64_t* a, uint64_t* out, size_t n) {                           \
    ZC_VEC_PROLOGUE(ctx, n, a && out);                                                                                    \
    half_kernel<MOD><<<grid_for(n), TPB, 0, ctx->stream>>>(a, out, n);                                                    \
    ZC_VEC_LAUNCHED(ctx);                                                                                                 \
  }                                                                                                                       \
  int32_t zc_##name##_half_batch(zc_ctx* ctx, const uint64_t* a, uint64_t* out, size_t n) {                               \
    ZC_VEC_PROLOGUE(ctx, n, a && out);                                                                                    \
    return host_run(ctx, a, n * 40, nullptr, 0, out, n * 40, nullptr, 0, [&](void* d0, void*, void* dout, void*) {        \
      return zc_##name##_half_batch_dev(ctx, (const uint64_t*)d0, (uint64_t*)dout, n);                                    \
    });                                                                                                                   \
  }                                                                                                                       \
  int32_t zc_##name##_to_bytes_batch_dev(zc_ctx* ctx, const uint64_t* a, uint8_t* out_bytes, size_t n) {                  \
    ZC_VEC_PROLOGUE(ctx, n, a && out_bytes);                                                                              \
    if ((uintptr_t)out_bytes & 3u) return zc_fail(ctx, ZC_ERR_SIZE, "out_bytes must be 4-byte aligned");                  \
    to_bytes_kernel<<<grid_for(n), TPB, 0, ctx->stream>>>(a, out_bytes, n);                                               \
    ZC_VEC_LAUNCHED(ctx);                                                                                                 \
  }                                                                                                                       \
  int32_t zc_##name##_to_bytes_batch(zc_ctx* ctx, const uint64_t* a, uint8_t* out_bytes, size_t n) {                      \
    ZC_VEC_PROLOGUE(ctx, n, a && out_bytes);                                                                              \
    return host_run(ctx, a, n * 40, nullptr, 0, out_bytes, n * 32, nullptr, 0, [&](void* d0, void*, void* dout, void*) {  \
      return zc_##name##_to_bytes_batch_dev(ctx, (const uint64_t*)d0, (uint8_t*)dout, n);                                 \
    });                                                                                                                   \
  }

ZC_DEFINE_POW_HALF(fe, ModP)
ZC_DEFINE_POW_HALF(scalar, ModL)
#undef ZC_DEFINE_POW_HALF

int32_t zc_fe_from_bytes_batch_dev(zc_ctx* ctx, const uint8_t* in_bytes, uint64_t* out, size_t n) {
  ZC_VEC_PROLOGUE(ctx, n, in_bytes && out);
  if ((uintptr_t)in_bytes & 3u) return zc_fail(ctx, ZC_ERR_SIZE, "in_bytes must be 4-byte aligned");
  from_bytes_kernel<false><<<grid_for(n), TPB, 0, ctx->stream>>>(in_bytes, out, nullptr, n);
  ZC_VEC_LAUNCHED(ctx);
}
int32_t zc_fe_from_bytes_batch(zc_ctx* ctx, const uint8_t* in_bytes, uint64_t* out, size_t n) {
  ZC_VEC_PROLOGUE(ctx, n, in_bytes && out);
  return host_run(ctx, in_bytes, n * 32, nullptr, 0, out, n * 40, nullptr, 0, [&](void* d0, void*, void* dout, void*) {
    return zc_fe_from_bytes_batch_dev(ctx, (const uint8_t*)d0, (uint64_t*)dout, n);
  });
}
int32_t zc_scalar_from_bytes_batch_dev(zc_ctx* ctx, const uint8_t* in_bytes, uint64_t* out, uint8_t* ok, size_t n) {
  ZC_VEC_PROLOGUE(ctx, n, in_bytes && out && ok);
  if ((uintptr_t)in_bytes & 3u) return zc_fail(ctx, ZC_ERR_SIZE, "in_bytes must be 4-byte aligned");
  from_bytes_kernel<true><<<grid_for(n), TPB, 0, ctx->stream>>>(in_bytes, out, ok, n);
  ZC_VEC_LAUNCHED(ctx);
}
int32_t zc_scalar_from_bytes_batch(zc_ctx* ctx, const uint8_t* in_bytes, uint64_t* out, uint8_t* ok, size_t n) {
  ZC_VEC_PROLOGUE(ctx, n, in_bytes && out && ok);
  return host_run(ctx, in_bytes, n * 32, nullptr, 0, out, n * 40, ok, n, [&](void* d0, void*, void* dout, void* dok) {
    return zc_scalar_from_bytes_batch_dev(ctx, (const uint8_t*)d0, (uint64_t*)dout, (uint8_t*)dok, n);
  });
}

int32_t zc_scalar_window_naf_batch_dev(zc_ctx* ctx, const uint64_t* a, int32_t width, int8_t* out_digits, size_t n) {
  ZC_VEC_PROLOGUE(ctx, n, a && out_digits);
  if (width < 2 || width > 7) return zc_fail(ctx, ZC_ERR_MODE, "NAF width must be in 2..7 (digits are i8)");
  if ((uintptr_t)out_digits & 3u) return zc_fail(ctx, ZC_ERR_SIZE, "out_digits must be 4-byte aligned");
  window_naf_kernel<<<grid_for(n), TPB, 0, ctx->stream>>>(a, width, out_digits, n);
  ZC_VEC_LAUNCHED(ctx);
}
int32_t zc_scalar_window_naf_batch(zc_ctx* ctx, const uint64_t* a, int32_t width, int8_t* out_digits, size_t n) {
  ZC_VEC_PROLOGUE(ctx, n, a && out_digits);
  if (width < 2 || width > 7) return zc_fail(ctx, ZC_ERR_MODE, "NAF width must be in 2..7 (digits are i8)");
  return host_run(ctx, a, n * 40, nullptr, 0, out_digits, n * 256, nullptr, 0, [&](void* d0, void*, void* dout, void*) {
    return zc_scalar_window_naf_batch_dev(ctx, (const uint64_t*)d0, width, (int8_t*)dout, n);
  });
}


int32_t zc_scalar_into_bits_batch_dev(zc_ctx* ctx, const uint64_t* a, uint8_t* out_bits, size_t n) {
  ZC_VEC_PROLOGUE(ctx, n, a && out_bits);
  if (((uintptr_t)out_bits & 15) != 0) return zc_fail(ctx, ZC_ERR_SIZE, "out_bits must be 16-byte aligned");
  into_bits_kernel<<<grid_for(16 * n), TPB, 0, ctx->stream>>>(a, out_bits, n);
  ZC_VEC_LAUNCHED(ctx);
}
int32_t zc_scalar_into_bits_batch(zc_ctx* ctx, const uint64_t* a, uint8_t* out_bits, size_t n) {
  ZC_VEC_PROLOGUE(ctx, n, a && out_bits);
  return host_run(ctx, a, n * 40, nullptr, 0, out_bits, n * 256, nullptr, 0, [&](void* d0, void*, void* dout, void*) {
    return zc_scalar_into_bits_batch_dev(ctx, (const uint64_t*)d0, (uint8_t*)dout, n);
  });
}

}  // extern "C"
