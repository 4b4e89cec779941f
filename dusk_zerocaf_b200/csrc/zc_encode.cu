// zc_encode.cu -- batched canonicalisation of points: the wire-format step right after the hot path.
//
//   zc_fe_invert_batch          FieldElement::inverse               /root/reference/src/backend/u64/field.rs:854-925
//   zc_point_to_affine_batch    AffinePoint::from(EdwardsPoint)     /root/reference/src/edwards.rs:1085-1092
//   zc_ristretto_compress_batch RistrettoPoint::compress            /root/reference/src/ristretto.rs:398-425
//                               (inv_sqrt / sqrt_ratio_i field.rs:443-503, is_positive :552-557, to_bytes :591-631)
// and the step right before it (8f rank 2):
//   zc_ristretto_decompress_batch CompressedRistretto::decompress   /root/reference/src/ristretto.rs:96-154
//   zc_point_is_valid_batch       ValidityCheck for EdwardsPoint    /root/reference/src/edwards.rs:393-400, 733-748
// and hash-to-group (8f rank 3):
//   zc_ristretto_elligator_batch          elligator_ristretto_flavor /root/reference/src/ristretto.rs:430-471
//   zc_ristretto_from_uniform_bytes_batch from_uniform_bytes         /root/reference/src/ristretto.rs:493-507
//
// The reference computes these with data-dependent loops (Savas-Koc almost-Montgomery inverse, Tonelli-Shanks); every
// one of them returns a uniquely defined VALUE (the inverse; the non-negative root), so the device code evaluates the
// same values with fixed exponent chains -- a^(p-2) and the p = 5 (mod 8) square-root-ratio recipe with exponent
// (p-5)/8 -- which keep all 32 lanes of a warp in lockstep.  Outputs are bit-identical to the reference's.
#include "zc_internal.h"
#include "zc_point.cuh"

using namespace zc;

namespace {

constexpr int TPB = 128;

// exponents as little-endian 32-bit words (uniform across threads: plain square-and-multiply, no divergence)
__constant__ uint32_t E_INV[8]  = {0x5cf5d3ebu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0u, 0u, 0u, 0x10000000u};   // p - 2      (253 bits)
__constant__ uint32_t E_SQRT[8] = {0x4b9eba7du, 0xcb024c63u, 0xd45ef39au, 0x029bdf3bu, 0u, 0u, 0u, 0x02000000u};   // (p - 5)/8  (250 bits)

__device__ __forceinline__ Fe SQRT_M1_MONT()      { return Fe{{0xa8c6c570u, 0xdb9954e7u, 0xfb49700du, 0xc85212d7u, 0x6de88652u, 0x28c7dc33u, 0x76c0a9d5u, 0x0aa666a6u}}; }   // constants.rs:96-102
__device__ __forceinline__ Fe INVSQRT_AMD_MONT()  { return Fe{{0x9c6a4e67u, 0xf2915c70u, 0x5445ce4du, 0x1200721bu, 0x86e004e5u, 0x4500326au, 0x2a9217d4u, 0x05b18190u}}; }   // constants.rs:123-129
__device__ __forceinline__ Fe MINUS_ONE_MONT()    { return Fe{{0xcf5d3ed0u, 0x812631a5u, 0x2f79cd65u, 0x4def9deau, 0x00000001u, 0u, 0u, 0u}}; }
__device__ __forceinline__ Fe POS_RANGE()         { return Fe{{0x2e7ae9f6u, 0x2c09318du, 0x517bce6bu, 0x0a6f7cefu, 0u, 0u, 0u, 0x08000000u}}; }   // (p-1)/2, constants.rs:12-13

// Out of line: the exponent loops call these ~315 times; one copy keeps the kernels small.
__device__ __noinline__ Fe mul_ni(Fe a, Fe b) { return mont_mul<ModP>(a, b); }
__device__ __noinline__ Fe sqr_ni(Fe a) { return mont_sqr<ModP>(a); }      // 36 + 32 wide multiplies instead of 64 + 32

// a^e for a Montgomery-form a and a constant exponent whose top set bit is bit nbits-1
__device__ __forceinline__ Fe fe_pow_const(const Fe& a, const uint32_t* __restrict__ e, int nbits) {
  Fe r = a;
#pragma unroll 1
  for (int bit = nbits - 2; bit >= 0; bit--) {
    r = mont_sqr<ModP>(r);                                   // inlined: ~250 of the ~315 products of a chain, no call marshalling
    if ((e[bit >> 5] >> (bit & 31)) & 1u) r = mul_ni(r, a);
  }
  return r;
}

// x <= (p-1)/2 for a canonical NORMAL-form x                       field.rs:552-557
__device__ __forceinline__ bool fe_is_positive(const Fe& x) {
  const Fe pr = POS_RANGE();
  uint32_t bw;
  asm("sub.cc.u32  %0, %1, %9;\n\t"
      "subc.cc.u32 %0, %2, %10;\n\t"
      "subc.cc.u32 %0, %3, %11;\n\t"
      "subc.cc.u32 %0, %4, %12;\n\t"
      "subc.cc.u32 %0, %5, %13;\n\t"
      "subc.cc.u32 %0, %6, %14;\n\t"
      "subc.cc.u32 %0, %7, %15;\n\t"
      "subc.cc.u32 %0, %8, %16;\n\t"
      "subc.u32    %0, 0, 0;\n\t"
      : "=&r"(bw)
      : "r"(pr.w[0]), "r"(pr.w[1]), "r"(pr.w[2]), "r"(pr.w[3]), "r"(pr.w[4]), "r"(pr.w[5]), "r"(pr.w[6]), "r"(pr.w[7]),
        "r"(x.w[0]), "r"(x.w[1]), "r"(x.w[2]), "r"(x.w[3]), "r"(x.w[4]), "r"(x.w[5]), "r"(x.w[6]), "r"(x.w[7]));
  return bw == 0;            // no borrow: POS_RANGE - x >= 0
}
// positivity of a Montgomery-form value (the predicate is defined on the value itself)
__device__ __forceinline__ bool mont_is_positive(const Fe& xm) { return fe_is_positive(from_mont<ModP>(xm)); }

// (was_square, +sqrt(1/v)) or (0, +sqrt(i/v)); v = 0 -> (0, 0).  Montgomery form in and out.     field.rs:443-503
__device__ __forceinline__ Fe mont_inv_sqrt(const Fe& v, bool& was_square) {
  typedef ModP M;
  const Fe one = Consts<M>::R1(), i = SQRT_M1_MONT();
  Fe v2 = sqr_ni(v);
  Fe v3 = mul_ni(v2, v);
  Fe v7 = mul_ni(sqr_ni(v3), v);
  Fe r = mul_ni(v3, fe_pow_const(v7, E_SQRT, 250));
  Fe check = mul_ni(v, sqr_ni(r));
  // check is one of 1, -1 (r <- i r), -i (r <- i r, non-square), i (non-square); 0 when v = 0 (r = 0 already)
  const bool minus_one = fe_eq(check, MINUS_ONE_MONT());
  was_square = minus_one || fe_eq(check, one);
  const bool flip = minus_one || fe_eq(check, fe_neg<M>(i));
  Fe ri = mul_ni(r, i);
  if (flip) r = ri;
  if (!mont_is_positive(r)) r = fe_neg<M>(r);
  return r;
}

// ---- inversion-based ops: Montgomery's trick, INV_K elements per thread -------------------------------------------------
// The inverse is a uniquely defined value, so any way of computing it is bit-identical to the reference's Savas-Koc loop
// (field.rs:854-925).  One a^(p-2) chain costs 251 squarings + 65 products; a thread that owns K elements multiplies them
// together (K products), inverts the product ONCE and peels the individual inverses off again (2 K products): 3 + 318 / K
// products per element instead of 318 -- 2^20 inversions 4.05 ms -> see profiles/.  Element j of thread t is index
// t + j T (T = threads in the grid), so every load / store of a warp stays coalesced and a thread only ever touches its own
// elements (in-place calls stay safe).  Zeros (inverse(0) = 0 here; the reference panics, field.rs:864) are replaced by one
// in the product and written back as zero.
// Representation trick: a loaded normal-form value a IS the Montgomery form of a / R, so the chain runs on the loaded words
// directly (no to_mont per element); what comes out is R^2 / a, and the factor is taken off the running inverse once per
// thread: SCALE = 2 Montgomery products by 1 give 1 / a (normal form), SCALE = 1 gives R / a (ready to multiply a normal-form
// value: a x (R / b) -> a / b).
constexpr int INV_K = 8;
template <int SCALE, class Load, class Store>
__device__ __forceinline__ void batch_invert(size_t n, Load&& load, Store&& store) {
  typedef ModP M;
  const size_t T = (size_t)gridDim.x * blockDim.x, t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const Fe one = Consts<M>::R1();
  Fe pref[INV_K];
  Fe acc = one;
#pragma unroll
  for (int j = 0; j < INV_K; j++) {
    const size_t i = t + (size_t)j * T;
    pref[j] = acc;
    if (i < n) {
      const Fe x = load(i);
      if (!fe_is_zero(x)) acc = mul_ni(acc, x);
    }
  }
  Fe inv = fe_pow_const(acc, E_INV, 253);
  const Fe unit{{1, 0, 0, 0, 0, 0, 0, 0}};
#pragma unroll
  for (int k = 0; k < SCALE; k++) inv = mul_ni(inv, unit);
#pragma unroll
  for (int j = INV_K - 1; j >= 0; j--) {
    const size_t i = t + (size_t)j * T;
    if (i < n) {
      const Fe x = load(i);
      const bool nz = !fe_is_zero(x);
      const Fe r = mul_ni(inv, pref[j]);
      if (nz) inv = mul_ni(inv, x);
      store(i, nz ? r : Fe{{0, 0, 0, 0, 0, 0, 0, 0}});
    }
  }
}
inline unsigned grid_inv(size_t n) { return (unsigned)(((n + INV_K - 1) / INV_K + TPB - 1) / TPB); }

__global__ void __launch_bounds__(TPB) fe_invert_kernel(const uint64_t* __restrict__ a, uint64_t* __restrict__ out, size_t n) {
  batch_invert<2>(n, [&](size_t i) { return fe_load52(a + 5 * i); }, [&](size_t i, const Fe& r) { fe_store52(out + 5 * i, r); });
}

// a / b = a * b^-1  (Div, field.rs:277-299; the reference asserts b != 0 -- here a / 0 = 0 like inverse(0))
__global__ void __launch_bounds__(TPB) fe_div_kernel(const uint64_t* __restrict__ a, const uint64_t* __restrict__ b, uint64_t* __restrict__ out, size_t n) {
  batch_invert<1>(n, [&](size_t i) { return fe_load52(b + 5 * i); },
                  [&](size_t i, const Fe& binv) { fe_store52(out + 5 * i, mul_ni(fe_load52(a + 5 * i), binv)); });   // normal * (R / b) -> normal
}

// (x, y) = (X / Z, Y / Z)                                                                        edwards.rs:1085-1092
__global__ void __launch_bounds__(TPB) pt_to_affine_kernel(const uint64_t* __restrict__ p, uint64_t* __restrict__ out, size_t n) {
  batch_invert<1>(n, [&](size_t i) { return fe_load52(p + 20 * i + 10); },
                  [&](size_t i, const Fe& zinv) {
                    fe_store52(out + 10 * i, mul_ni(fe_load52(p + 20 * i), zinv));
                    fe_store52(out + 10 * i + 5, mul_ni(fe_load52(p + 20 * i + 5), zinv));
                  });
}

// Ristretto encoding                                                                             ristretto.rs:398-425
__global__ void __launch_bounds__(TPB) ristretto_compress_kernel(const uint64_t* __restrict__ p, uint8_t* __restrict__ out, size_t n) {
  typedef ModP M;
  size_t idx = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (idx >= n) return;
  Pt P = pt_to_mont(pt_load52(p + 20 * idx));
  const Fe i = SQRT_M1_MONT();
  Fe u1 = mul_ni(fe_add<M>(P.Z, P.Y), fe_sub<M>(P.Z, P.Y));
  Fe u2 = mul_ni(P.X, P.Y);
  bool sq;
  Fe I = mont_inv_sqrt(mul_ni(u1, sqr_ni(u2)), sq);
  Fe D1 = mul_ni(u1, I);
  Fe D2 = mul_ni(u2, I);
  Fe Zinv = mul_ni(mul_ni(D1, D2), P.T);
  Fe x = P.X, y = P.Y, D = D2;
  if (!mont_is_positive(mul_ni(P.T, Zinv))) {
    x = mul_ni(i, P.Y);
    y = mul_ni(i, P.X);
    D = mul_ni(D1, INVSQRT_AMD_MONT());
  }
  if (!mont_is_positive(mul_ni(x, Zinv))) y = fe_neg<M>(y);
  Fe s = from_mont<M>(mul_ni(fe_sub<M>(P.Z, y), D));
  if (!fe_is_positive(s)) s = fe_neg<M>(s);
  // to_bytes: the canonical value, little-endian (field.rs:591-631)
  uint4* o = reinterpret_cast<uint4*>(out + 32 * idx);
  o[0] = make_uint4(s.w[0], s.w[1], s.w[2], s.w[3]);
  o[1] = make_uint4(s.w[4], s.w[5], s.w[6], s.w[7]);
}

// Ristretto decoding: ok[i] = 1 and out[i] = (x, y, 1, x y), or ok[i] = 0 and out[i] = 0            ristretto.rs:96-154
// from_bytes keeps all 256 bits (field.rs:563-587), so the reference's re-encoding check always passes and step 1 reduces
// to is_positive: the encoded integer must be <= (p-1)/2.
__global__ void __launch_bounds__(TPB) ristretto_decompress_kernel(const uint8_t* __restrict__ in, uint64_t* __restrict__ out,
                                                                   uint8_t* __restrict__ ok, size_t n) {
  typedef ModP M;
  size_t idx = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (idx >= n) return;
  const uint32_t* w = reinterpret_cast<const uint32_t*>(in + 32 * idx);
  Fe sn{{w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]}};
  bool good = fe_is_positive(sn);
  if (!good) sn = Fe{{0, 0, 0, 0, 0, 0, 0, 0}};            // keep the arithmetic below on canonical input
  const Fe one = Consts<M>::R1();
  Fe s = to_mont<M>(sn);
  Fe ss = sqr_ni(s);
  Fe u1 = fe_sub<M>(one, ss);                              // 1 + a s^2, a = -1
  Fe u2 = fe_add<M>(one, ss);
  Fe u2_sq = sqr_ni(u2);
  Fe v = fe_sub<M>(fe_neg<M>(mul_ni(D_MONT(), sqr_ni(u1))), u2_sq);
  bool sq;
  Fe I = mont_inv_sqrt(mul_ni(v, u2_sq), sq);
  good = good && sq;
  Fe Dx = mul_ni(I, u2);
  Fe Dy = mul_ni(mul_ni(I, Dx), v);
  Fe x = mul_ni(fe_add<M>(s, s), Dx);
  if (!mont_is_positive(x)) x = fe_neg<M>(x);
  Fe y = mul_ni(u1, Dy);
  Fe t = mul_ni(x, y);
  good = good && mont_is_positive(t) && !fe_is_zero(y);
  uint64_t* o = out + 20 * idx;
  if (good) {
    fe_store52(o, from_mont<M>(x));
    fe_store52(o + 5, from_mont<M>(y));
    o[10] = 1; o[11] = 0; o[12] = 0; o[13] = 0; o[14] = 0;
    fe_store52(o + 15, from_mont<M>(t));
  } else {
#pragma unroll
    for (int k = 0; k < 20; k++) o[k] = 0;
  }
  ok[idx] = good ? 1 : 0;
}

// (a X^2 + Y^2) Z^2 == Z^4 + d X^2 Y^2                                                        edwards.rs:733-748
__global__ void __launch_bounds__(TPB) pt_is_valid_kernel(const uint64_t* __restrict__ p, uint8_t* __restrict__ ok, size_t n) {
  typedef ModP M;
  size_t idx = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (idx >= n) return;
  Fe X = to_mont<M>(fe_load52(p + 20 * idx)), Y = to_mont<M>(fe_load52(p + 20 * idx + 5)), Z = to_mont<M>(fe_load52(p + 20 * idx + 10));
  Fe xx = sqr_ni(X), yy = sqr_ni(Y), zz = sqr_ni(Z);
  Fe left = mul_ni(fe_sub<M>(yy, xx), zz);
  Fe right = fe_add<M>(sqr_ni(zz), mul_ni(mul_ni(D_MONT(), xx), yy));
  ok[idx] = fe_eq(left, right) ? 1 : 0;
}

// ---- hash to group -------------------------------------------------------------------------------------------------
__device__ __forceinline__ Fe ONE_MINUS_D_SQ_MONT()    { return Fe{{0xdf568a00u, 0x939bafb4u, 0xcb683c61u, 0xcf54cbd6u, 0xf592c91cu, 0xec300527u, 0x6551a020u, 0x08f2ce8eu}}; }
__device__ __forceinline__ Fe D_MINUS_ONE_SQ_MONT()    { return Fe{{0x46acfacfu, 0xd6c8ed3du, 0x910abc15u, 0xa96623f1u, 0x2515d67au, 0x410f7878u, 0x702e8474u, 0x0f7142b0u}}; }
__device__ __forceinline__ Fe SQRT_AD_MINUS_ONE_MONT() { return Fe{{0x9e882cc7u, 0xdf3921e3u, 0x98d6d78eu, 0xbba9f682u, 0xa71a2909u, 0xe38d903eu, 0x70c5212du, 0x030de965u}}; }   // constants.rs:132-138

// (was_square, +sqrt(u/v)) or (0, +sqrt(i u/v)); u = 0 -> (1, 0); v = 0 != u -> (0, 0).          field.rs:461-503
__device__ __forceinline__ Fe mont_sqrt_ratio_i(const Fe& u, const Fe& v, bool& was_square) {
  typedef ModP M;
  const Fe i = SQRT_M1_MONT();
  Fe v2 = sqr_ni(v);
  Fe v3 = mul_ni(v2, v);
  Fe v7 = mul_ni(sqr_ni(v3), v);
  Fe r = mul_ni(mul_ni(u, v3), fe_pow_const(mul_ni(u, v7), E_SQRT, 250));
  Fe check = mul_ni(v, sqr_ni(r));
  const Fe neg_u = fe_neg<M>(u);
  const Fe neg_ui = mul_ni(neg_u, i);
  const bool correct = fe_eq(check, u), flipped = fe_eq(check, neg_u), flipped_i = fe_eq(check, neg_ui);
  Fe ri = mul_ni(r, i);
  if (flipped || flipped_i) r = ri;
  if (!mont_is_positive(r)) r = fe_neg<M>(r);
  was_square = correct || flipped;
  return r;
}

// (was_square, r) with r the non-negative root of u/v, or of i u/v when u/v is not a square; (1, 0) for u = 0, (0, 0) for
// v = 0, u != 0                                                                                      field.rs:443-491
__global__ void __launch_bounds__(TPB) fe_sqrt_ratio_i_kernel(const uint64_t* __restrict__ u, const uint64_t* __restrict__ v,
                                                              uint64_t* __restrict__ out, uint8_t* __restrict__ was_square, size_t n) {
  typedef ModP M;
  size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  bool sq;
  Fe r = mont_sqrt_ratio_i(to_mont<M>(fe_load52(u + 5 * i)), to_mont<M>(fe_load52(v + 5 * i)), sq);
  fe_store52(out + 5 * i, from_mont<M>(r));
  was_square[i] = sq ? 1 : 0;
}

// Ristretto-flavoured Elligator 2 on a Montgomery-form r0; the returned (X:Y:Z:T) are the reference's own products
__device__ __forceinline__ Pt elligator_mont(const Fe& r0) {
  typedef ModP M;
  const Fe one = Consts<M>::R1(), d = D_MONT();
  Fe c = MINUS_ONE_MONT();
  Fe r = mul_ni(SQRT_M1_MONT(), sqr_ni(r0));
  Fe Ns = mul_ni(fe_add<M>(r, one), ONE_MINUS_D_SQ_MONT());
  Fe D = mul_ni(fe_sub<M>(c, mul_ni(d, r)), fe_add<M>(r, d));
  bool sq;
  Fe s = mont_sqrt_ratio_i(Ns, D, sq);
  Fe s_prim = mul_ni(s, r0);
  if (mont_is_positive(s_prim)) s_prim = fe_neg<M>(s_prim);       // s' = -|s r0|
  if (!sq) { s = s_prim; c = r; }
  Fe Nt = fe_sub<M>(mul_ni(mul_ni(c, fe_sub<M>(r, one)), D_MINUS_ONE_SQ_MONT()), D);
  Fe ss = sqr_ni(s);
  Fe W0 = mul_ni(fe_add<M>(s, s), D);
  Fe W1 = mul_ni(Nt, SQRT_AD_MINUS_ONE_MONT());
  Fe W2 = fe_sub<M>(one, ss);
  Fe W3 = fe_add<M>(one, ss);
  return Pt{mul_ni(W0, W3), mul_ni(W2, W1), mul_ni(W1, W3), mul_ni(W0, W2)};
}

// any 256-bit integer -> Montgomery form of its residue (x R^2 < R m, so one product reduces it)
__device__ __forceinline__ Fe bytes_to_mont(const uint8_t* __restrict__ b) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(b);
  return to_mont<ModP>(Fe{{w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]}});
}

__global__ void __launch_bounds__(TPB) ristretto_elligator_kernel(const uint64_t* __restrict__ r0, uint64_t* __restrict__ out, size_t n) {
  size_t idx = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (idx >= n) return;
  pt_store52(out + 20 * idx, pt_from_mont(elligator_mont(to_mont<ModP>(fe_load52(r0 + 5 * idx)))));
}

__global__ void __launch_bounds__(TPB) ristretto_from_uniform_kernel(const uint8_t* __restrict__ in, uint64_t* __restrict__ out, size_t n) {
  size_t idx = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (idx >= n) return;
  Pt R1 = elligator_mont(bytes_to_mont(in + 64 * idx));
  Pt R2 = elligator_mont(bytes_to_mont(in + 64 * idx + 32));
  pt_store52(out + 20 * idx, pt_from_mont(pt_add_ref(R1, R2)));      // R_1 + R_2 with the reference Add
}

inline unsigned grid_for(size_t n) { return (unsigned)((n + TPB - 1) / TPB); }
constexpr size_t MAX_N = (size_t)1 << 31;

// small synchronous host wrapper (these are not bandwidth-bound: ~320 multiplications per element)
template <class F>
int32_t host_unary(zc_ctx* ctx, const void* in, size_t in_bytes, void* out, size_t out_bytes, F run) {
  void *din = nullptr, *dout = nullptr;
  int32_t rc;
  if ((rc = zc_scratch(ctx, 0, in_bytes, &din))) return rc;
  if ((rc = zc_scratch(ctx, 2, out_bytes, &dout))) return rc;
  ZC_CUDA(ctx, cudaMemcpyAsync(din, in, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = run(din, dout))) return rc;
  ZC_CUDA(ctx, cudaMemcpyAsync(out, dout, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  ZC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZC_OK;
}

}  // namespace

#define ZC_ENC_PROLOGUE(ctx, n, cond)                                                           \
  do {                                                                                          \
    if (!(ctx)) return ZC_ERR_NULL;                                                             \
    ZC_CUDA(ctx, cudaSetDevice((ctx)->device));                                                 \
    if ((n) > MAX_N) return zc_fail(ctx, ZC_ERR_SIZE, "n exceeds 2^31");                        \
    if ((n) == 0) return ZC_OK;                                                                 \
    if (!(cond)) return zc_fail(ctx, ZC_ERR_NULL, "null pointer argument");                     \
  } while (0)

extern "C" {

int32_t zc_fe_invert_batch_dev(zc_ctx* ctx, const uint64_t* a, uint64_t* out, size_t n) {
  ZC_ENC_PROLOGUE(ctx, n, a && out);
  fe_invert_kernel<<<grid_inv(n), TPB, 0, ctx->stream>>>(a, out, n);
  ctx->launches++;
  ZC_CUDA(ctx, cudaGetLastError());
  return ZC_OK;
}
int32_t zc_fe_invert_batch(zc_ctx* ctx, const uint64_t* a, uint64_t* out, size_t n) {
  ZC_ENC_PROLOGUE(ctx, n, a && out);
  return host_unary(ctx, a, n * 40, out, n * 40, [&](void* di, void* dout) {
    return zc_fe_invert_batch_dev(ctx, (const uint64_t*)di, (uint64_t*)dout, n);
  });
}

int32_t zc_fe_div_batch_dev(zc_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  ZC_ENC_PROLOGUE(ctx, n, a && b && out);
  fe_div_kernel<<<grid_inv(n), TPB, 0, ctx->stream>>>(a, b, out, n);
  ctx->launches++;
  ZC_CUDA(ctx, cudaGetLastError());
  return ZC_OK;
}
int32_t zc_fe_div_batch(zc_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  ZC_ENC_PROLOGUE(ctx, n, a && b && out);
  void *da = nullptr, *db = nullptr, *dout = nullptr;
  int32_t rc;
  if ((rc = zc_scratch(ctx, 0, n * 40, &da))) return rc;
  if ((rc = zc_scratch(ctx, 1, n * 40, &db))) return rc;
  if ((rc = zc_scratch(ctx, 2, n * 40, &dout))) return rc;
  ZC_CUDA(ctx, cudaMemcpyAsync(da, a, n * 40, cudaMemcpyHostToDevice, ctx->stream));
  ZC_CUDA(ctx, cudaMemcpyAsync(db, b, n * 40, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = zc_fe_div_batch_dev(ctx, (const uint64_t*)da, (const uint64_t*)db, (uint64_t*)dout, n))) return rc;
  ZC_CUDA(ctx, cudaMemcpyAsync(out, dout, n * 40, cudaMemcpyDeviceToHost, ctx->stream));
  ZC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZC_OK;
}

int32_t zc_fe_sqrt_ratio_i_batch_dev(zc_ctx* ctx, const uint64_t* u, const uint64_t* v, uint64_t* out, uint8_t* was_square, size_t n) {
  ZC_ENC_PROLOGUE(ctx, n, u && v && out && was_square);
  fe_sqrt_ratio_i_kernel<<<grid_for(n), TPB, 0, ctx->stream>>>(u, v, out, was_square, n);
  ctx->launches++;
  ZC_CUDA(ctx, cudaGetLastError());
  return ZC_OK;
}
int32_t zc_fe_sqrt_ratio_i_batch(zc_ctx* ctx, const uint64_t* u, const uint64_t* v, uint64_t* out, uint8_t* was_square, size_t n) {
  ZC_ENC_PROLOGUE(ctx, n, u && v && out && was_square);
  void *du = nullptr, *dv = nullptr, *dout = nullptr, *dsq = nullptr;
  int32_t rc;
  if ((rc = zc_scratch(ctx, 0, n * 40, &du))) return rc;
  if ((rc = zc_scratch(ctx, 1, n * 40, &dv))) return rc;
  if ((rc = zc_scratch(ctx, 2, n * 40, &dout))) return rc;
  if ((rc = zc_scratch(ctx, 3, n, &dsq))) return rc;
  ZC_CUDA(ctx, cudaMemcpyAsync(du, u, n * 40, cudaMemcpyHostToDevice, ctx->stream));
  ZC_CUDA(ctx, cudaMemcpyAsync(dv, v, n * 40, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = zc_fe_sqrt_ratio_i_batch_dev(ctx, (const uint64_t*)du, (const uint64_t*)dv, (uint64_t*)dout, (uint8_t*)dsq, n))) return rc;
  ZC_CUDA(ctx, cudaMemcpyAsync(out, dout, n * 40, cudaMemcpyDeviceToHost, ctx->stream));
  ZC_CUDA(ctx, cudaMemcpyAsync(was_square, dsq, n, cudaMemcpyDeviceToHost, ctx->stream));
  ZC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZC_OK;
}

int32_t zc_point_to_affine_batch_dev(zc_ctx* ctx, const uint64_t* p, uint64_t* out_xy, size_t n) {
  ZC_ENC_PROLOGUE(ctx, n, p && out_xy);
  pt_to_affine_kernel<<<grid_inv(n), TPB, 0, ctx->stream>>>(p, out_xy, n);
  ctx->launches++;
  ZC_CUDA(ctx, cudaGetLastError());
  return ZC_OK;
}
int32_t zc_point_to_affine_batch(zc_ctx* ctx, const uint64_t* p, uint64_t* out_xy, size_t n) {
  ZC_ENC_PROLOGUE(ctx, n, p && out_xy);
  return host_unary(ctx, p, n * 160, out_xy, n * 80, [&](void* di, void* dout) {
    return zc_point_to_affine_batch_dev(ctx, (const uint64_t*)di, (uint64_t*)dout, n);
  });
}

int32_t zc_ristretto_compress_batch_dev(zc_ctx* ctx, const uint64_t* p, uint8_t* out_bytes, size_t n) {
  ZC_ENC_PROLOGUE(ctx, n, p && out_bytes);
  if ((uintptr_t)out_bytes & 15u) return zc_fail(ctx, ZC_ERR_SIZE, "out_bytes must be 16-byte aligned");
  ristretto_compress_kernel<<<grid_for(n), TPB, 0, ctx->stream>>>(p, out_bytes, n);
  ctx->launches++;
  ZC_CUDA(ctx, cudaGetLastError());
  return ZC_OK;
}
int32_t zc_ristretto_compress_batch(zc_ctx* ctx, const uint64_t* p, uint8_t* out_bytes, size_t n) {
  ZC_ENC_PROLOGUE(ctx, n, p && out_bytes);
  return host_unary(ctx, p, n * 160, out_bytes, n * 32, [&](void* di, void* dout) {
    return zc_ristretto_compress_batch_dev(ctx, (const uint64_t*)di, (uint8_t*)dout, n);
  });
}

int32_t zc_ristretto_decompress_batch_dev(zc_ctx* ctx, const uint8_t* in_bytes, uint64_t* out_points, uint8_t* ok, size_t n) {
  ZC_ENC_PROLOGUE(ctx, n, in_bytes && out_points && ok);
  if ((uintptr_t)in_bytes & 3u) return zc_fail(ctx, ZC_ERR_SIZE, "in_bytes must be 4-byte aligned");
  ristretto_decompress_kernel<<<grid_for(n), TPB, 0, ctx->stream>>>(in_bytes, out_points, ok, n);
  ctx->launches++;
  ZC_CUDA(ctx, cudaGetLastError());
  return ZC_OK;
}
int32_t zc_ristretto_decompress_batch(zc_ctx* ctx, const uint8_t* in_bytes, uint64_t* out_points, uint8_t* ok, size_t n) {
  ZC_ENC_PROLOGUE(ctx, n, in_bytes && out_points && ok);
  void *din = nullptr, *dout = nullptr, *dok = nullptr;
  int32_t rc;
  if ((rc = zc_scratch(ctx, 0, n * 32, &din))) return rc;
  if ((rc = zc_scratch(ctx, 2, n * 160, &dout))) return rc;
  if ((rc = zc_scratch(ctx, 1, n, &dok))) return rc;
  ZC_CUDA(ctx, cudaMemcpyAsync(din, in_bytes, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = zc_ristretto_decompress_batch_dev(ctx, (const uint8_t*)din, (uint64_t*)dout, (uint8_t*)dok, n))) return rc;
  ZC_CUDA(ctx, cudaMemcpyAsync(out_points, dout, n * 160, cudaMemcpyDeviceToHost, ctx->stream));
  ZC_CUDA(ctx, cudaMemcpyAsync(ok, dok, n, cudaMemcpyDeviceToHost, ctx->stream));
  ZC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZC_OK;
}

int32_t zc_point_is_valid_batch_dev(zc_ctx* ctx, const uint64_t* p, uint8_t* ok, size_t n) {
  ZC_ENC_PROLOGUE(ctx, n, p && ok);
  pt_is_valid_kernel<<<grid_for(n), TPB, 0, ctx->stream>>>(p, ok, n);
  ctx->launches++;
  ZC_CUDA(ctx, cudaGetLastError());
  return ZC_OK;
}
int32_t zc_point_is_valid_batch(zc_ctx* ctx, const uint64_t* p, uint8_t* ok, size_t n) {
  ZC_ENC_PROLOGUE(ctx, n, p && ok);
  return host_unary(ctx, p, n * 160, ok, n, [&](void* di, void* dout) {
    return zc_point_is_valid_batch_dev(ctx, (const uint64_t*)di, (uint8_t*)dout, n);
  });
}

int32_t zc_ristretto_elligator_batch_dev(zc_ctx* ctx, const uint64_t* r0, uint64_t* out_points, size_t n) {
  ZC_ENC_PROLOGUE(ctx, n, r0 && out_points);
  ristretto_elligator_kernel<<<grid_for(n), TPB, 0, ctx->stream>>>(r0, out_points, n);
  ctx->launches++;
  ZC_CUDA(ctx, cudaGetLastError());
  return ZC_OK;
}
int32_t zc_ristretto_elligator_batch(zc_ctx* ctx, const uint64_t* r0, uint64_t* out_points, size_t n) {
  ZC_ENC_PROLOGUE(ctx, n, r0 && out_points);
  return host_unary(ctx, r0, n * 40, out_points, n * 160, [&](void* di, void* dout) {
    return zc_ristretto_elligator_batch_dev(ctx, (const uint64_t*)di, (uint64_t*)dout, n);
  });
}

int32_t zc_ristretto_from_uniform_bytes_batch_dev(zc_ctx* ctx, const uint8_t* in_bytes, uint64_t* out_points, size_t n) {
  ZC_ENC_PROLOGUE(ctx, n, in_bytes && out_points);
  if ((uintptr_t)in_bytes & 3u) return zc_fail(ctx, ZC_ERR_SIZE, "in_bytes must be 4-byte aligned");
  ristretto_from_uniform_kernel<<<grid_for(n), TPB, 0, ctx->stream>>>(in_bytes, out_points, n);
  ctx->launches++;
  ZC_CUDA(ctx, cudaGetLastError());
  return ZC_OK;
}
int32_t zc_ristretto_from_uniform_bytes_batch(zc_ctx* ctx, const uint8_t* in_bytes, uint64_t* out_points, size_t n) {
  ZC_ENC_PROLOGUE(ctx, n, in_bytes && out_points);
  return host_unary(ctx, in_bytes, n * 64, out_points, n * 160, [&](void* di, void* dout) {
    return zc_ristretto_from_uniform_bytes_batch_dev(ctx, (const uint8_t*)di, (uint64_t*)dout, n);
  });
}

}  // extern "C"
