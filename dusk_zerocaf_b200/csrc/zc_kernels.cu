// zc_kernels.cu -- batched field / scalar / point kernels and their C-ABI entry points (include/zerocaf_b200.h).
//
// One residue (or one point) per thread; operands stream through HBM in the reference's own AoS radix-2^52 layout and
// are repacked to 8 x u32 Montgomery words in registers (zc_fe.cuh).  Every kernel is memory-streaming or
// integer-multiply bound; none uses tensor cores (there is no dense contraction on this path).
#include <stdlib.h>

#include "zc_internal.h"
#include "zc_point.cuh"

using namespace zc;

namespace {

constexpr int TPB = 256;

enum FeOp { OP_MUL = 0, OP_SQUARE = 1, OP_ADD = 2, OP_SUB = 3, OP_NEG = 4 };

// ---- K1: out[i] = a[i] (op) b[i]   (field.rs:191-315 / scalar.rs:184-283) -------------------------------
// Normal-form operands: full product + two folds with 2^K = -c (mod m)  (fe_mul_normal_pre, zc_fe.cuh) -- no Montgomery.
template <class M, int OP>
__global__ void __launch_bounds__(TPB) fe_op_kernel(const uint64_t* __restrict__ a, const uint64_t* __restrict__ b,
                                                    uint64_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  Fe r;
  if (OP == OP_MUL) {
    r = fe_mul_normal_pre<M>(fe_load52_shl<Shape<M>::SA>(a + 5 * i), fe_load52_shl<Shape<M>::SB>(b + 5 * i));
    fe_store52(out + 5 * i, r);
    return;
  }
  if (OP == OP_SQUARE) {
    fe_store52(out + 5 * i, fe_sqr_normal_pre<M>(fe_load52_shl<Shape<M>::SA>(a + 5 * i)));
    return;
  }
  Fe x = fe_load52(a + 5 * i);
  if (OP == OP_ADD) {
    Fe y = fe_load52(b + 5 * i);
    r = fe_add<M>(x, y);
  } else if (OP == OP_SUB) {
    Fe y = fe_load52(b + 5 * i);
    r = fe_sub<M>(x, y);
  } else {
    r = fe_neg<M>(x);
  }
  fe_store52(out + 5 * i, r);
}

// ---- K1 fused (BASELINE config 2): prod = a*b, sq = a^2 from one load of a
template <class M>
__global__ void __launch_bounds__(TPB) fe_mul_square_kernel(const uint64_t* __restrict__ a, const uint64_t* __restrict__ b,
                                                            uint64_t* __restrict__ prod, uint64_t* __restrict__ sq, size_t n) {
  size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  // operands unpacked straight into their pre-shifted form (SA for a -- shared by the product and the square -- SB for b)
  const Fe x = fe_load52_shl<Shape<M>::SA>(a + 5 * i);
  const Fe y = fe_load52_shl<Shape<M>::SB>(b + 5 * i);
  fe_store52(prod + 5 * i, fe_mul_normal_pre<M>(x, y));
  fe_store52(sq + 5 * i, fe_sqr_normal_pre<M>(x));
}

// ---- K1 fused on the 32-byte wire format (FieldElement::to_bytes / from_bytes, field.rs:563-631): the little-endian
// encoding of a canonical value IS its eight 32-bit words, so an element is two 16-byte loads and the kernel moves the
// algorithmic 128 bytes per pair (the [u64;5] limb layout moves 160).  Same values, bit for bit.
template <int S>
__device__ __forceinline__ Fe fe_load32_shl(const uint8_t* __restrict__ p) {
  const uint4 lo = *reinterpret_cast<const uint4*>(p), hi = *reinterpret_cast<const uint4*>(p + 16);
  const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
  Fe r;
  r.w[0] = w[0] << S;
#pragma unroll
  for (int k = 1; k < 8; k++) r.w[k] = __funnelshift_l(w[k - 1], w[k], S);
  return r;
}
__device__ __forceinline__ void fe_store32(uint8_t* __restrict__ p, const Fe& a) {
  *reinterpret_cast<uint4*>(p) = make_uint4(a.w[0], a.w[1], a.w[2], a.w[3]);
  *reinterpret_cast<uint4*>(p + 16) = make_uint4(a.w[4], a.w[5], a.w[6], a.w[7]);
}
template <class M>
__global__ void __launch_bounds__(TPB) fe_mul_square_packed_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                                   uint8_t* __restrict__ prod, uint8_t* __restrict__ sq, size_t n) {
  size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  const Fe x = fe_load32_shl<Shape<M>::SA>(a + 32 * i);
  const Fe y = fe_load32_shl<Shape<M>::SB>(b + 32 * i);
  fe_store32(prod + 32 * i, fe_mul_normal_pre<M>(x, y));
  fe_store32(sq + 32 * i, fe_sqr_normal_pre<M>(x));
}

// ---- K2: point add / sub / double / neg on the ABI layout, limb-exact (edwards.rs:440-592) ----------------
enum PtOp { PT_ADD = 0, PT_SUB = 1, PT_DOUBLE = 2, PT_NEG = 3 };

// pt_add_ref_normal (zc_point.cuh) with a selectable mix of inlined products and calls of one out-of-line multiplier:
// ZC_PT_INLINE bit k = product k of (A, B, TT, C, D, E, F', H', X3, Y3, Z3, T3) inlined.  Twelve inlined products are ~44 KB of
// straight-line code, more than the 32 KB instruction cache; every call costs ~16 IMAD.MOV of marshalling.
// Measured at 2^22 additions in 128 x 4 blocks (profiles/r02_pt_add_inline_ab.txt): 0xfff (all inlined, round 1) 0.833 ms,
// 0x000 0.770, 0xf00 0.771, 0xf03 0.756, 0xfc0 0.754, 0xfc3 0.740, 0xf0f 0.745, 0x3c0 0.753, 0xfcf 0.793, 0xfe0 0.738,
// 0xff0 (default: the first four products through the call, eight inlined) 0.739 ms.
#ifndef ZC_PT_INLINE
#define ZC_PT_INLINE 0xff0
#endif
__device__ __noinline__ Fe pt_mul_ni(Fe a, Fe b) { return mont_mul<ModP>(a, b); }
#define ZC_PT_MUL(K, X, Y) (((ZC_PT_INLINE >> (K)) & 1) ? mont_mul<ModP>((X), (Y)) : pt_mul_ni((X), (Y)))
__device__ __forceinline__ Pt pt_add_ref_normal_mix(const Pt& p, const Pt& q) {
  typedef ModP M;
  Fe A = ZC_PT_MUL(0, p.X, q.X);
  Fe B = ZC_PT_MUL(1, p.Y, q.Y);
  Fe C = ZC_PT_MUL(3, ZC_PT_MUL(2, p.T, q.T), D_MONT());
  Fe D = ZC_PT_MUL(4, p.Z, q.Z);
  Fe E = ZC_PT_MUL(5, fe_add<M>(p.X, p.Y), fe_add<M>(q.X, q.Y));
  E = fe_sub<M>(fe_sub<M>(E, A), B);
  Fe F = fe_sub<M>(D, C);
  Fe G = fe_add<M>(D, C);
  Fe H = fe_add<M>(B, A);
  const Fe R4 = Consts<M>::R4();
  F = ZC_PT_MUL(6, F, R4);
  H = ZC_PT_MUL(7, H, R4);
  Pt r;
  r.X = ZC_PT_MUL(8, E, F);
  r.Y = ZC_PT_MUL(9, G, H);
  r.Z = ZC_PT_MUL(10, F, G);
  r.T = ZC_PT_MUL(11, E, H);
  return r;
}
#undef ZC_PT_MUL

// Block shape: BT threads, at least MINB resident blocks per SM (the register cap that goes with it).  ZC_PT_VARIANT = 0: 256 x 2,
// 1 (default): 128 x 4 -- with the mixed inlined / out-of-line products below this is the fastest (0.739 ms per 2^22 additions;
// 128 x 5 needs 96 registers and spills around the calls: 0.81 ms), 2: 128 x 5, 3: 128 x 6.
template <int OP, int BT, int MINB>
__global__ void __launch_bounds__(BT, MINB) pt_op_kernel(const uint64_t* __restrict__ p, const uint64_t* __restrict__ q,
                                                         uint64_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * BT + threadIdx.x;
  if (i >= n) return;
  Pt a = pt_load52(p + 20 * i);
  Pt r;
  if (OP == PT_ADD) {
    Pt b = pt_load52(q + 20 * i);
    r = pt_add_ref_normal_mix(a, b);
  } else if (OP == PT_SUB) {
    Pt b = pt_neg(pt_load52(q + 20 * i));   // Sub = self + (-other), edwards.rs:512
    r = pt_add_ref_normal_mix(a, b);
  } else if (OP == PT_DOUBLE) {
    r = pt_add_ref_normal_mix(a, a);        // Double = self + self, edwards.rs:589-591
  } else {
    r = pt_neg(a);
  }
  pt_store52(out + 20 * i, r);
}

// ---- Ristretto equality (ristretto.rs:166-176) -----------------------------------------------------------
__global__ void __launch_bounds__(TPB) ristretto_eq_kernel(const uint64_t* __restrict__ p, const uint64_t* __restrict__ q,
                                                           uint8_t* __restrict__ eq, size_t n) {
  size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  typedef ModP M;
  Fe X1 = fe_load52(p + 20 * i), Y1 = fe_load52(p + 20 * i + 5);
  Fe X2 = fe_load52(q + 20 * i), Y2 = fe_load52(q + 20 * i + 5);
  // a common factor 1/R on both sides does not change equality
  bool e1 = fe_eq(mont_mul<M>(X1, Y2), mont_mul<M>(Y1, X2));
  bool e2 = fe_eq(mont_mul<M>(X1, X2), mont_mul<M>(Y1, Y2));
  eq[i] = (e1 || e2) ? 1 : 0;
}

// ---- K3 strict: [s]P with the reference's LSB-first double_and_add (edwards.rs:102-120), limb-exact ----------
// Every lane must perform exactly the reference's additions (Q += N on set bits, N += N every bit) in the reference's
// order, but different lanes have different bits.  Executing "Q += N" for the whole warp whenever ANY lane has the bit set
// costs 2 additions per bit (498 per scalar instead of the ~374 a lane needs).  Instead each lane parks the N it has to
// add later in a small ring in shared memory and the warp runs a "drain" step -- every lane with a parked value does one
// Q += value -- only when some lane's ring is full (or the scalars are exhausted): ~173 drain steps instead of 249 for
// a ring of 3.  The order of a lane's own Q additions is unchanged, so the result is limb for limb the reference's.
// The loop body is ONE inlined addition whose right operand is always read from the ring (the doubling step first
// stores N into the ring's write slot, which is also how N gets parked when the bit is set).
constexpr int STRICT_TPB = 128;
constexpr int STRICT_RING = 4;            // 3 parked values + the write slot
__device__ __forceinline__ Fe ring_ld(const uint4* __restrict__ q) {
  const uint4 lo = q[0], hi = q[STRICT_TPB];
  return Fe{{lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w}};
}
__device__ __forceinline__ void ring_st(uint4* __restrict__ q, const Fe& f) {
  q[0] = make_uint4(f.w[0], f.w[1], f.w[2], f.w[3]);
  q[STRICT_TPB] = make_uint4(f.w[4], f.w[5], f.w[6], f.w[7]);
}
// p + q with the reference's formulas (pt_add_ref), q = (X, Y, Z, T) staged in shared memory at pieces 0-1, 2-3, 4-5, 6-7
__device__ __forceinline__ Pt pt_add_ref_staged(const Pt& p, const uint4* __restrict__ q) {
  typedef ModP M;
  const Fe qX = ring_ld(q), qY = ring_ld(q + 2 * STRICT_TPB);
  Fe A = mont_mul<M>(p.X, qX);
  Fe B = mont_mul<M>(p.Y, qY);
  Fe E = mont_mul<M>(fe_add<M>(p.X, p.Y), fe_add<M>(qX, qY));
  E = fe_sub<M>(fe_sub<M>(E, A), B);
  Fe C = mont_mul<M>(mont_mul<M>(p.T, ring_ld(q + 6 * STRICT_TPB)), D_MONT());
  Fe D = mont_mul<M>(p.Z, ring_ld(q + 4 * STRICT_TPB));
  Fe F = fe_sub<M>(D, C);
  Fe G = fe_add<M>(D, C);
  Fe H = fe_add<M>(B, A);
  Pt r;
  r.X = mont_mul<M>(E, F);
  r.Y = mont_mul<M>(G, H);
  r.Z = mont_mul<M>(F, G);
  r.T = mont_mul<M>(E, H);
  return r;
}

__global__ void __launch_bounds__(STRICT_TPB) scalar_mul_strict_kernel(const uint64_t* __restrict__ points,
                                                                       const uint64_t* __restrict__ scalars,
                                                                       uint64_t* __restrict__ out, size_t n) {
  extern __shared__ uint4 ring[];          // [slot][16-byte piece 0..7][thread]
  const int tx = threadIdx.x;
  size_t i = (size_t)blockIdx.x * STRICT_TPB + tx;
  const bool live = i < n;
  size_t ii = live ? i : 0;
  Pt N = pt_to_mont(pt_load52(points + 20 * ii));
  Fe s = fe_load52(scalars + 5 * ii);
  if (!live) { for (int k = 0; k < 8; k++) s.w[k] = 0; }
  Pt Q = pt_identity_mont();
  int head = 0, tail = 0, count = 0;       // ring slots head .. head+count-1 are parked; slot tail is the write slot
  for (;;) {
    const bool nz = !fe_is_zero(s);
    const bool odd = nz && (s.w[0] & 1u) != 0;
    const bool any_nz = __any_sync(0xffffffffu, nz);
    if (!any_nz && !__any_sync(0xffffffffu, count > 0)) break;
    // drain when a lane that has to park N has no room, or when only parked values are left
    const bool drain = !any_nz || __any_sync(0xffffffffu, odd && count == STRICT_RING - 1);
    bool active;
    int slot;
    Pt lhs;
    if (drain) {
      active = count > 0; slot = head; lhs = Q;
    } else {
      active = nz; slot = tail; lhs = N;
      uint4* w = ring + (size_t)slot * 8 * STRICT_TPB + tx;
      ring_st(w, N.X); ring_st(w + 2 * STRICT_TPB, N.Y); ring_st(w + 4 * STRICT_TPB, N.Z); ring_st(w + 6 * STRICT_TPB, N.T);
    }
    const Pt r = pt_add_ref_staged(lhs, ring + (size_t)slot * 8 * STRICT_TPB + tx);
    if (drain) {
      if (active) { Q = r; head = (head + 1 == STRICT_RING) ? 0 : head + 1; count--; }
    } else if (active) {
      N = r;                                             // N = N.double() = N + N   edwards.rs:589-591
      if (odd) { tail = (tail + 1 == STRICT_RING) ? 0 : tail + 1; count++; }   // park the pre-doubling N for Q += N
      // n = n.half_without_mod()   scalar.rs:562-574
#pragma unroll
      for (int k = 0; k < 7; k++) s.w[k] = (s.w[k] >> 1) | (s.w[k + 1] << 31);
      s.w[7] >>= 1;
    }
  }
  if (live) pt_store52(out + 20 * i, pt_from_mont(Q));
}

// ---- K3 fast: signed 4-bit fixed window, dedicated doubling, per-thread table of cached multiples in global scratch ----
// digits d_j in [-8, 8), s = sum d_j 16^j; the table holds 1P..8P in cached form (Y+X, Y-X, Z, 2dT); 63 windows cover
// 252 bits (s < L < 2^250).  The table (8 x 128 B per point) lives in a grow-only scratch arena indexed by the thread's
// slot in the (persistent) grid, laid out [entry][16-byte word][slot] so that every table access of a warp is one
// coalesced 512-byte transaction and mostly an L2 hit.  Keeping it out of shared memory leaves the kernel limited by
// registers only (4 blocks of 128 threads per SM instead of 3 blocks of 64), which is what the dependent multiply chains
// need to keep the integer pipe busy.
constexpr int SM_FAST_TPB = 128;
// One out-of-line multiplier keeps the window loop (a doubling and an addition, 16 products) at a few KB of code instead
// of ~60 KB of inlined carry chains: with 16 warps per SM spread over the loop, instruction fetch was a visible stall.
__device__ __noinline__ Fe smf_mul(Fe a, Fe b) { return mont_mul<ModP>(a, b); }
__device__ __noinline__ Fe smf_sqr(Fe a) { return mont_sqr<ModP>(a); }     // dedicated squaring: 36 + 32 wide multiplies
// need_t == false: the caller's next operation is a doubling, which reads X, Y, Z only -- T3 = E H is skipped (the T of the
// returned point is then stale).  Three of the four doublings of a window and every window addition but the last run that
// way: 36 instead of 40 products per window.  need_t is uniform over the grid (a loop counter), never a divergent branch.
__device__ __forceinline__ Pt smf_double(const Pt& p, bool need_t) {           // pt_double_fast (dbl-2008-hwcd, a = -1)
  typedef ModP M;
#ifndef ZC_SMF_INLINE
#define ZC_SMF_INLINE 1      // measured at 2^20: 0 -> 36.93 ms, 1 -> 36.36 ms, 3 -> 38.04 ms, 4 -> 36.93 ms, 5 -> 39.20 ms (instruction cache)
#endif
#define ZC_SMF_SQR(X) (((ZC_SMF_INLINE >> 2) & 1) ? mont_sqr<ModP>(X) : smf_sqr(X))      // bit 2: the doubling's squarings inlined
  Fe A = ZC_SMF_SQR(p.X), B = ZC_SMF_SQR(p.Y), Z2 = ZC_SMF_SQR(p.Z);
  Fe C = fe_add<M>(Z2, Z2);
  Fe D = fe_neg<M>(A);
  Fe S = fe_add<M>(p.X, p.Y);
  Fe E = fe_sub<M>(fe_sub<M>(ZC_SMF_SQR(S), A), B);
  Fe G = fe_add<M>(D, B);
  Fe F = fe_sub<M>(G, C);
  Fe H = fe_sub<M>(D, B);
  // ZC_SMF_INLINE bit 0: the doubling's output products inlined (no register marshalling around the calls), bit 1: the addition's
#define ZC_SMF_OUT(BIT, X, Y) (((ZC_SMF_INLINE >> (BIT)) & 1) ? mont_mul<ModP>((X), (Y)) : smf_mul((X), (Y)))
  Pt r{ZC_SMF_OUT(0, E, F), ZC_SMF_OUT(0, G, H), ZC_SMF_OUT(0, F, G), p.T};
  if (need_t) r.T = ZC_SMF_OUT(0, E, H);
  return r;
}
__device__ __forceinline__ Pt smf_add(const Pt& p, const PtCached& q, bool need_t) {   // pt_add_cached (add-2008-hwcd-3)
  typedef ModP M;
  Fe A = smf_mul(fe_sub<M>(p.Y, p.X), q.YmX);
  Fe B = smf_mul(fe_add<M>(p.Y, p.X), q.YpX);
  Fe C = smf_mul(p.T, q.T2d);
  Fe D = smf_mul(p.Z, q.Z);
  D = fe_add<M>(D, D);
  Fe E = fe_sub<M>(B, A), F = fe_sub<M>(D, C), G = fe_add<M>(D, C), H = fe_add<M>(B, A);
  Pt r{ZC_SMF_OUT(1, E, F), ZC_SMF_OUT(1, G, H), ZC_SMF_OUT(1, F, G), p.T};
  if (need_t) r.T = ZC_SMF_OUT(1, E, H);
  return r;
}
// SMEM_TABLE (experiment, ZC_SMF_SMEM=1): the per-thread table in shared memory instead of the global arena.  1 KiB per thread
// means ONE 128-thread CTA per SM (128 KiB) -- 4 warps per SM instead of 16; measured 2^20 scalar multiplications in
// profiles/r02_fixed_base_tma_ab.txt.  The product keeps the coalesced global arena (L2-resident, registers-limited occupancy).
template <bool SMEM_TABLE>
__global__ void __launch_bounds__(SM_FAST_TPB, SMEM_TABLE ? 1 : 4) scalar_mul_fast_kernel(const uint64_t* __restrict__ points,
                                                                         const uint64_t* __restrict__ scalars,
                                                                         uint64_t* __restrict__ out, size_t n,
                                                                         uint4* __restrict__ table_g, size_t nslots_g) {
  extern __shared__ __align__(16) uint4 smf_smem[];
  const size_t slot_g = (size_t)blockIdx.x * SM_FAST_TPB + threadIdx.x;
  uint4* const table = SMEM_TABLE ? smf_smem : table_g;
  const size_t nslots = SMEM_TABLE ? (size_t)SM_FAST_TPB : nslots_g;             // table stride
  const size_t slot = SMEM_TABLE ? (size_t)threadIdx.x : slot_g;                 // table slot
  auto store_entry = [&](int e, const PtCached& c) {
    uint4* base = table + (size_t)e * 8 * nslots + slot;
    base[0 * nslots] = make_uint4(c.YpX.w[0], c.YpX.w[1], c.YpX.w[2], c.YpX.w[3]); base[1 * nslots] = make_uint4(c.YpX.w[4], c.YpX.w[5], c.YpX.w[6], c.YpX.w[7]);
    base[2 * nslots] = make_uint4(c.YmX.w[0], c.YmX.w[1], c.YmX.w[2], c.YmX.w[3]); base[3 * nslots] = make_uint4(c.YmX.w[4], c.YmX.w[5], c.YmX.w[6], c.YmX.w[7]);
    base[4 * nslots] = make_uint4(c.Z.w[0], c.Z.w[1], c.Z.w[2], c.Z.w[3]);         base[5 * nslots] = make_uint4(c.Z.w[4], c.Z.w[5], c.Z.w[6], c.Z.w[7]);
    base[6 * nslots] = make_uint4(c.T2d.w[0], c.T2d.w[1], c.T2d.w[2], c.T2d.w[3]); base[7 * nslots] = make_uint4(c.T2d.w[4], c.T2d.w[5], c.T2d.w[6], c.T2d.w[7]);
  };
  auto load_fe = [&](const uint4* p, Fe& f) {
    uint4 lo = p[0], hi = p[nslots];
    f.w[0] = lo.x; f.w[1] = lo.y; f.w[2] = lo.z; f.w[3] = lo.w; f.w[4] = hi.x; f.w[5] = hi.y; f.w[6] = hi.z; f.w[7] = hi.w;
  };
  auto load_entry = [&](int e) {
    PtCached c;
    const uint4* base = table + (size_t)e * 8 * nslots + slot;
    load_fe(base, c.YpX); load_fe(base + 2 * nslots, c.YmX); load_fe(base + 4 * nslots, c.Z); load_fe(base + 6 * nslots, c.T2d);
    return c;
  };
  for (size_t i = slot_g; ; i += nslots_g) {
    // whole warps leave together (the loop body uses warp votes)
    if (!__any_sync(0xffffffffu, i < n)) break;
    const bool live = i < n;
    const size_t ii = live ? i : 0;
    Pt P = pt_to_mont(pt_load52(points + 20 * ii));
    Fe s = fe_load52(scalars + 5 * ii);
    {  // table: e -> (e+1) P
      PtCached c1 = pt_to_cached(P);
      store_entry(0, c1);
      Pt acc = P;
#pragma unroll 1
      for (int e = 1; e < 8; e++) {
        acc = smf_add(acc, c1, true);
        store_entry(e, pt_to_cached(acc));
      }
    }
    // signed digits: recode from the bottom into a packed 4-bit array (two's-complement nibbles)
    uint32_t dig[8];
    {
      uint32_t carry = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        uint32_t w = s.w[k], o = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          uint32_t d = ((w >> (4 * j)) & 15u) + carry;   // 0..16
          carry = (d >= 8u) ? 1u : 0u;                   // d in [8,16] -> d - 16, carry 1
          o |= (d & 15u) << (4 * j);
        }
        dig[k] = o;
      }
      // s < 2^250 so the top nibble (bits 252..255) is 0 before the carry and the final carry is always 0
    }
    Pt Q = pt_identity_mont();
#pragma unroll 1
    for (int j = 63; j >= 0; j--) {
      if (j != 63) {
#pragma unroll 1
        for (int r = 0; r < 4; r++) Q = smf_double(Q, r == 3);       // only the doubling in front of the addition needs T
      }
      uint32_t nib = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) if ((j >> 3) == k) nib = dig[k];
      nib = (nib >> (4 * (j & 7))) & 15u;
      const int d = (nib >= 8u) ? (int)nib - 16 : (int)nib;
      const int mag = d < 0 ? -d : d;
      if (__any_sync(0xffffffffu, mag != 0)) {
        PtCached c = load_entry(mag ? mag - 1 : 0);
        if (d < 0) c = pt_cached_neg(c);
        Pt r = smf_add(Q, c, j == 0);                                  // doublings follow unless this is the last window
        if (mag != 0) Q = r;
      }
    }
    if (live) pt_store52(out + 20 * i, pt_from_mont(Q));
  }
}

// ---- fold k points in index order with the reference Add (one thread; k is the number of ranks) ---------------
__global__ void point_fold_kernel(const uint64_t* __restrict__ pts, size_t k, uint64_t* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (k == 0) { pt_store52(out, pt_from_mont(pt_identity_mont())); return; }
  Pt acc = pt_to_mont(pt_load52(pts));
  for (size_t j = 1; j < k; j++) acc = pt_add_ref(acc, pt_to_mont(pt_load52(pts + 20 * j)));
  pt_store52(out, pt_from_mont(acc));
}

inline unsigned grid_for(size_t n, int tpb) { return (unsigned)((n + tpb - 1) / tpb); }

// ---- input validation (opt-in: zc_ctx_set_validation, zc_*_check_canonical_batch) ---------------------------------
// The reference's types expose their limbs (`pub [u64;5]`), and its own tests build non-canonical values on purpose
// (field.rs:1160-1167, 1193-1200: operands with FIELD_L added).  The kernels here assume canonical inputs (limbs < 2^52,
// value < modulus): with validation on, every element of every input array is checked on the device and the first
// offending element index (array order: a, then b) comes back as ZC_ERR_NONCANONICAL instead of a silently wrong result.
template <class M>
__global__ void __launch_bounds__(TPB) check_canonical_kernel(const uint64_t* __restrict__ a, size_t n, unsigned long long base,
                                                              unsigned long long* __restrict__ first_bad) {
  size_t i = (size_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  const uint64_t* p = a + 5 * i;
  const uint64_t l0 = p[0], l1 = p[1], l2 = p[2], l3 = p[3], l4 = p[4];
  bool bad = ((l0 | l1 | l2 | l3) >> 52) != 0 || (l4 >> 48) != 0;      // 4 x 52 + 48 = 256 bits
  if (!bad) {
    Fe x = fe_from_limbs52(l0, l1, l2, l3, l4);
    Fe y = x;
    reduce_once<M>(y);                                                   // y != x  <=>  x >= m
    bad = !fe_eq(x, y);
  }
  if (bad) atomicMin(first_bad, base + (unsigned long long)i);
}

int32_t validation_setup(zc_ctx* ctx) {
  if (ctx->vflag_dev) return ZC_OK;
  ZC_CUDA(ctx, cudaMalloc(&ctx->vflag_dev, 8));
  ZC_CUDA(ctx, cudaMemsetAsync(ctx->vflag_dev, 0xff, 8, ctx->stream));
  return ZC_OK;
}
// enqueue the check of n_elems residues (a point is four) on the context's stream
template <class M>
int32_t validation_enqueue(zc_ctx* ctx, const uint64_t* a, size_t n_elems, size_t base) {
  int32_t rc;
  if ((rc = validation_setup(ctx))) return rc;
  if (n_elems == 0 || !a) return ZC_OK;
  check_canonical_kernel<M><<<grid_for(n_elems, TPB), TPB, 0, ctx->stream>>>(a, n_elems, (unsigned long long)base, ctx->vflag_dev);
  ctx->launches++;
  ZC_CUDA(ctx, cudaGetLastError());
  return ZC_OK;
}
// wait, read the verdict, re-arm.  elems_per_unit: 1 (field / scalar arrays) or 4 (points) for the index in the message
int32_t validation_finish(zc_ctx* ctx, int elems_per_unit, unsigned long long* first_bad_out) {
  if (!ctx->vflag_dev) return ZC_OK;
  unsigned long long v = 0;
  ZC_CUDA(ctx, cudaMemcpyAsync(&v, ctx->vflag_dev, 8, cudaMemcpyDeviceToHost, ctx->stream));
  ZC_CUDA(ctx, cudaMemsetAsync(ctx->vflag_dev, 0xff, 8, ctx->stream));
  ZC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (v == ~0ull) return ZC_OK;
  if (first_bad_out) *first_bad_out = v / (unsigned long long)elems_per_unit;
  snprintf(ctx->err, sizeof(ctx->err), "non-canonical input (limb >= 2^52 or value >= modulus): first offending element %llu",
           v / (unsigned long long)elems_per_unit);
  return ZC_ERR_NONCANONICAL;
}

template <class M, int OP>
int32_t launch_fe(zc_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  if (n == 0) return ZC_OK;
  if (ctx->validate) {
    int32_t rc;
    if ((rc = validation_enqueue<M>(ctx, a, n, ctx->vbase))) return rc;
    if (b && (rc = validation_enqueue<M>(ctx, b, n, ctx->vbase))) return rc;
  }
  fe_op_kernel<M, OP><<<grid_for(n, TPB), TPB, 0, ctx->stream>>>(a, b, out, n);
  ctx->launches++;
  ZC_CUDA(ctx, cudaGetLastError());
  return ZC_OK;
}
template <int OP>
int32_t launch_pt(zc_ctx* ctx, const uint64_t* p, const uint64_t* q, uint64_t* out, size_t n) {
  if (n == 0) return ZC_OK;
  if (ctx->validate) {
    int32_t rc;
    if ((rc = validation_enqueue<ModP>(ctx, p, 4 * n, 4 * ctx->vbase))) return rc;
    if (q && (rc = validation_enqueue<ModP>(ctx, q, 4 * n, 4 * ctx->vbase))) return rc;
  }
  static const int variant = getenv("ZC_PT_VARIANT") ? atoi(getenv("ZC_PT_VARIANT")) : 1;
  if (variant == 1) pt_op_kernel<OP, 128, 4><<<grid_for(n, 128), 128, 0, ctx->stream>>>(p, q, out, n);
  else if (variant == 3) pt_op_kernel<OP, 128, 6><<<grid_for(n, 128), 128, 0, ctx->stream>>>(p, q, out, n);
  else if (variant != 0) pt_op_kernel<OP, 128, 5><<<grid_for(n, 128), 128, 0, ctx->stream>>>(p, q, out, n);
  else pt_op_kernel<OP, TPB, 2><<<grid_for(n, TPB), TPB, 0, ctx->stream>>>(p, q, out, n);
  ctx->launches++;
  ZC_CUDA(ctx, cudaGetLastError());
  return ZC_OK;
}

constexpr size_t MAX_N = (size_t)1 << 31;

// ---- host-pointer entry points: chunked, three-stage pipeline ----------------------------------------------------
// H2D of chunk k+1 (copy-in stream), the kernel on chunk k (the context's stream) and D2H of chunk k-1 (copy-out
// stream) overlap, so a PCIe-bound call costs max(H2D, D2H) instead of their sum.  Device staging is the grow-only
// scratch of the context (whole arrays, so chunks never alias).  Pageable host memory still works (the copies then
// serialise inside the driver); zc_host_alloc / zc_host_register give pinned memory and the full overlap.
struct HostArr { const void* h_in; void* h_out; size_t stride; int slot; void* d; };

constexpr size_t PIPE_CHUNK_BYTES = (size_t)16 << 20;

int32_t pipe_setup(zc_ctx* ctx) {
  if (ctx->copy_in) return ZC_OK;
  ZC_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
  ZC_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
  for (int i = 0; i < 2 * ZC_PIPE_MAX_CHUNKS; i++) ZC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->pipe_ev[i], cudaEventDisableTiming));
  return ZC_OK;
}

template <class F>
int32_t host_pipelined(zc_ctx* ctx, size_t n, HostArr* ins, int n_in, HostArr* outs, int n_out, F launch) {
  int32_t rc;
  if ((rc = pipe_setup(ctx))) return rc;
  size_t max_stride = 1;
  for (int i = 0; i < n_in; i++) { if ((rc = zc_scratch(ctx, ins[i].slot, n * ins[i].stride, &ins[i].d))) return rc; if (ins[i].stride > max_stride) max_stride = ins[i].stride; }
  for (int i = 0; i < n_out; i++) { if ((rc = zc_scratch(ctx, outs[i].slot, n * outs[i].stride, &outs[i].d))) return rc; if (outs[i].stride > max_stride) max_stride = outs[i].stride; }
  static const size_t chunk_bytes = getenv("ZC_PIPE_CHUNK_MB") ? (size_t)atoi(getenv("ZC_PIPE_CHUNK_MB")) << 20 : PIPE_CHUNK_BYTES;
  size_t chunk = chunk_bytes / max_stride;
  if (chunk >= 65536) chunk &= ~(size_t)65535;   // chunk boundaries on 64 Ki elements: page-aligned DMA (measured: 29.6 vs 32.5 ms)
  if (chunk < 4096) chunk = 4096;
  if ((n + chunk - 1) / chunk > ZC_PIPE_MAX_CHUNKS) chunk = (n + ZC_PIPE_MAX_CHUNKS - 1) / ZC_PIPE_MAX_CHUNKS;
  // the copy streams must not run ahead of work already queued on the context's stream
  ZC_CUDA(ctx, cudaEventRecord(ctx->pipe_ev[0], ctx->stream));
  ZC_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_in, ctx->pipe_ev[0], 0));
  int k = 0;
  for (size_t i0 = 0; i0 < n; i0 += chunk, k++) {
    const size_t cnt = (n - i0 < chunk) ? n - i0 : chunk;
    cudaEvent_t ev_in = ctx->pipe_ev[2 * k], ev_done = ctx->pipe_ev[2 * k + 1];
    for (int i = 0; i < n_in; i++)
      ZC_CUDA(ctx, cudaMemcpyAsync((char*)ins[i].d + i0 * ins[i].stride, (const char*)ins[i].h_in + i0 * ins[i].stride,
                                   cnt * ins[i].stride, cudaMemcpyHostToDevice, ctx->copy_in));
    ZC_CUDA(ctx, cudaEventRecord(ev_in, ctx->copy_in));
    ZC_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_in, 0));
    ctx->vbase = i0;
    ctx->in_pipeline = true;                      // _dev entry points called from a chunk do not read the verdict back (the caller does, once)
    rc = launch(i0, cnt);
    ctx->in_pipeline = false;
    ctx->vbase = 0;
    if (rc) return rc;
    ZC_CUDA(ctx, cudaEventRecord(ev_done, ctx->stream));
    ZC_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_out, ev_done, 0));
    for (int i = 0; i < n_out; i++)
      ZC_CUDA(ctx, cudaMemcpyAsync((char*)outs[i].h_out + i0 * outs[i].stride, (const char*)outs[i].d + i0 * outs[i].stride,
                                   cnt * outs[i].stride, cudaMemcpyDeviceToHost, ctx->copy_out));
  }
  ZC_CUDA(ctx, cudaStreamSynchronize(ctx->copy_out));
  ZC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZC_OK;
}

// two inputs (b may be null) -> one output
template <class F>
int32_t host_binary(zc_ctx* ctx, size_t n, const void* a, size_t a_stride, const void* b, size_t b_stride, void* out, size_t out_stride, F run) {
  HostArr ins[2] = {{a, nullptr, a_stride, 0, nullptr}, {b, nullptr, b_stride, 1, nullptr}};
  HostArr outs[1] = {{nullptr, out, out_stride, 2, nullptr}};
  return host_pipelined(ctx, n, ins, b ? 2 : 1, outs, 1, [&](size_t i0, size_t cnt) {
    return run((const char*)ins[0].d + i0 * a_stride, b ? (const char*)ins[1].d + i0 * b_stride : nullptr,
               (char*)outs[0].d + i0 * out_stride, cnt);
  });
}

}  // namespace

// ================================================= C ABI =================================================
#define ZC_CHECK_CTX(ctx) do { if (!(ctx)) return ZC_ERR_NULL; ZC_CUDA(ctx, cudaSetDevice((ctx)->device)); } while (0)
#define ZC_CHECK_PTR(ctx, cond) do { if (!(cond)) return zc_fail(ctx, ZC_ERR_NULL, "null pointer argument"); } while (0)
#define ZC_CHECK_N(ctx, n) do { if ((n) > MAX_N) return zc_fail(ctx, ZC_ERR_SIZE, "n exceeds 2^31"); } while (0)

extern "C" {

const char* zc_version(void) { return "zerocaf_b200 0.1 (sm_100a)"; }

int32_t zc_ctx_create(int32_t device, void* stream, zc_ctx** out) {
  if (!out) return ZC_ERR_NULL;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess) return -(int32_t)e;          // no CUDA device: there is no CPU fallback
  if (device < 0 || device >= count) return ZC_ERR_SIZE;
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return -(int32_t)e;
  zc_ctx* ctx = new zc_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
  if (stream) {
    ctx->stream = (cudaStream_t)stream;
  } else {
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete ctx; return -(int32_t)e; }
    ctx->own_stream = true;
  }
  cudaFuncSetAttribute(scalar_mul_strict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STRICT_RING * 8 * STRICT_TPB * 16);
  *out = ctx;
  return ZC_OK;
}

int32_t zc_ctx_destroy(zc_ctx* ctx) {
  if (!ctx) return ZC_ERR_NULL;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < 6; i++) if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
  if (ctx->msm_ws) cudaFree(ctx->msm_ws);
  if (ctx->basepoint_table) cudaFree(ctx->basepoint_table);
  if (ctx->gather_buf) cudaFree(ctx->gather_buf);
  if (ctx->peers_connected && ctx->peers_ipc) for (int r = 0; r < ctx->nranks; r++) if (r != ctx->rank && ctx->peers.p[r]) cudaIpcCloseMemHandle(ctx->peers.p[r]);
  if (ctx->mailbox) cudaFree(ctx->mailbox);
  if (ctx->vflag_dev) cudaFree(ctx->vflag_dev);
  if (ctx->peer_error_host) cudaFreeHost(ctx->peer_error_host);
  if (ctx->copy_in) { cudaStreamDestroy(ctx->copy_in); cudaStreamDestroy(ctx->copy_out); for (int i = 0; i < 2 * ZC_PIPE_MAX_CHUNKS; i++) if (ctx->pipe_ev[i]) cudaEventDestroy(ctx->pipe_ev[i]); }
  if (ctx->msm_graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)ctx->msm_graph_exec);
  if (ctx->side_stream) { cudaStreamSynchronize(ctx->side_stream); cudaStreamDestroy(ctx->side_stream); cudaStreamSynchronize(ctx->chain_stream); cudaStreamDestroy(ctx->chain_stream); for (int i = 0; i < 3; i++) { if (ctx->side_extra[i]) cudaStreamDestroy(ctx->side_extra[i]); if (ctx->side_hi[i]) cudaStreamDestroy(ctx->side_hi[i]); } if (ctx->sort_stream) cudaStreamDestroy(ctx->sort_stream); if (ctx->sort_hi) cudaStreamDestroy(ctx->sort_hi); if (ctx->acc2) cudaStreamDestroy(ctx->acc2); }
  for (int i = 0; i < 16; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return ZC_OK;
}

int32_t zc_ctx_sync(zc_ctx* ctx) {
  ZC_CHECK_CTX(ctx);
  ZC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return zc_peer_check_error(ctx);                  // a timed-out sharded exchange surfaces here (ZC_ERR_STATE), once
}

const char* zc_last_error_string(zc_ctx* ctx) { return ctx ? ctx->err : "null context"; }
uint64_t zc_ctx_launch_count(zc_ctx* ctx) { return ctx ? ctx->launches : 0; }

int32_t zc_host_alloc(size_t bytes, void** out) {
  if (!out) return ZC_ERR_NULL;
  cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
  return e == cudaSuccess ? ZC_OK : -(int32_t)e;
}
int32_t zc_host_free(void* p) {
  cudaError_t e = cudaFreeHost(p);
  return e == cudaSuccess ? ZC_OK : -(int32_t)e;
}
int32_t zc_host_register(void* p, size_t bytes) {
  if (!p) return ZC_ERR_NULL;
  cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
  return e == cudaSuccess ? ZC_OK : -(int32_t)e;
}
int32_t zc_host_unregister(void* p) {
  if (!p) return ZC_ERR_NULL;
  cudaError_t e = cudaHostUnregister(p);
  return e == cudaSuccess ? ZC_OK : -(int32_t)e;
}

// ---- field / scalar element-wise ------------------------------------------------------------------------
#define ZC_DEFINE_BIN(NAME, MOD, OP)                                                                              \
  int32_t NAME##_dev(zc_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {                \
    ZC_CHECK_CTX(ctx); ZC_CHECK_N(ctx, n);                                                                        \
    if (n == 0) return ZC_OK;                                                                                     \
    ZC_CHECK_PTR(ctx, a && b && out);                                                                             \
    int32_t rc_ = launch_fe<MOD, OP>(ctx, a, b, out, n);                                                          \
    return (rc_ == ZC_OK && ctx->validate) ? validation_finish(ctx, 1, nullptr) : rc_;                            \
  }                                                                                                               \
  int32_t NAME(zc_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {                      \
    ZC_CHECK_CTX(ctx); ZC_CHECK_N(ctx, n);                                                                        \
    if (n == 0) return ZC_OK;                                                                                     \
    ZC_CHECK_PTR(ctx, a && b && out);                                                                             \
    int32_t rc_ = host_binary(ctx, n, a, 40, b, 40, out, 40, [&](const void* da, const void* db, void* dout, size_t m) { \
      return launch_fe<MOD, OP>(ctx, (const uint64_t*)da, (const uint64_t*)db, (uint64_t*)dout, m);               \
    });                                                                                                           \
    return (rc_ == ZC_OK && ctx->validate) ? validation_finish(ctx, 1, nullptr) : rc_;                            \
  }
#define ZC_DEFINE_UN(NAME, MOD, OP)                                                                               \
  int32_t NAME##_dev(zc_ctx* ctx, const uint64_t* a, uint64_t* out, size_t n) {                                   \
    ZC_CHECK_CTX(ctx); ZC_CHECK_N(ctx, n);                                                                        \
    if (n == 0) return ZC_OK;                                                                                     \
    ZC_CHECK_PTR(ctx, a && out);                                                                                  \
    int32_t rc_ = launch_fe<MOD, OP>(ctx, a, nullptr, out, n);                                                    \
    return (rc_ == ZC_OK && ctx->validate) ? validation_finish(ctx, 1, nullptr) : rc_;                            \
  }                                                                                                               \
  int32_t NAME(zc_ctx* ctx, const uint64_t* a, uint64_t* out, size_t n) {                                         \
    ZC_CHECK_CTX(ctx); ZC_CHECK_N(ctx, n);                                                                        \
    if (n == 0) return ZC_OK;                                                                                     \
    ZC_CHECK_PTR(ctx, a && out);                                                                                  \
    int32_t rc_ = host_binary(ctx, n, a, 40, nullptr, 0, out, 40, [&](const void* da, const void*, void* dout, size_t m) { \
      return launch_fe<MOD, OP>(ctx, (const uint64_t*)da, nullptr, (uint64_t*)dout, m);                           \
    });                                                                                                           \
    return (rc_ == ZC_OK && ctx->validate) ? validation_finish(ctx, 1, nullptr) : rc_;                            \
  }

ZC_DEFINE_BIN(zc_fe_mul_batch, ModP, OP_MUL)
ZC_DEFINE_BIN(zc_fe_add_batch, ModP, OP_ADD)
ZC_DEFINE_BIN(zc_fe_sub_batch, ModP, OP_SUB)
ZC_DEFINE_UN(zc_fe_square_batch, ModP, OP_SQUARE)
ZC_DEFINE_UN(zc_fe_neg_batch, ModP, OP_NEG)
ZC_DEFINE_BIN(zc_scalar_mul_batch, ModL, OP_MUL)
ZC_DEFINE_BIN(zc_scalar_add_batch, ModL, OP_ADD)
ZC_DEFINE_BIN(zc_scalar_sub_batch, ModL, OP_SUB)
ZC_DEFINE_UN(zc_scalar_square_batch, ModL, OP_SQUARE)
ZC_DEFINE_UN(zc_scalar_neg_batch, ModL, OP_NEG)

int32_t zc_fe_mul_square_batch_dev(zc_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t* prod, uint64_t* sq, size_t n) {
  ZC_CHECK_CTX(ctx); ZC_CHECK_N(ctx, n);
  if (n == 0) return ZC_OK;
  ZC_CHECK_PTR(ctx, a && b && prod && sq);
  if (ctx->validate) {
    int32_t rc;
    if ((rc = validation_enqueue<ModP>(ctx, a, n, ctx->vbase))) return rc;
    if ((rc = validation_enqueue<ModP>(ctx, b, n, ctx->vbase))) return rc;
  }
  fe_mul_square_kernel<ModP><<<grid_for(n, TPB), TPB, 0, ctx->stream>>>(a, b, prod, sq, n);
  ctx->launches++;
  ZC_CUDA(ctx, cudaGetLastError());
  return (ctx->validate && !ctx->in_pipeline) ? validation_finish(ctx, 1, nullptr) : ZC_OK;
}
int32_t zc_fe_mul_square_batch(zc_ctx* ctx, const uint64_t* a, const uint64_t* b, uint64_t* prod, uint64_t* sq, size_t n) {
  ZC_CHECK_CTX(ctx); ZC_CHECK_N(ctx, n);
  if (n == 0) return ZC_OK;
  ZC_CHECK_PTR(ctx, a && b && prod && sq);
  HostArr ins[2] = {{a, nullptr, 40, 0, nullptr}, {b, nullptr, 40, 1, nullptr}};
  HostArr outs[2] = {{nullptr, prod, 40, 2, nullptr}, {nullptr, sq, 40, 3, nullptr}};
  int32_t rc_ = host_pipelined(ctx, n, ins, 2, outs, 2, [&](size_t i0, size_t cnt) {
    return zc_fe_mul_square_batch_dev(ctx, (const uint64_t*)ins[0].d + 5 * i0, (const uint64_t*)ins[1].d + 5 * i0,
                                      (uint64_t*)outs[0].d + 5 * i0, (uint64_t*)outs[1].d + 5 * i0, cnt);
  });
  return (rc_ == ZC_OK && ctx->validate) ? validation_finish(ctx, 1, nullptr) : rc_;
}

int32_t zc_fe_mul_square_batch_packed_dev(zc_ctx* ctx, const uint8_t* a, const uint8_t* b, uint8_t* prod, uint8_t* sq, size_t n) {
  ZC_CHECK_CTX(ctx); ZC_CHECK_N(ctx, n);
  if (n == 0) return ZC_OK;
  ZC_CHECK_PTR(ctx, a && b && prod && sq);
  if ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)prod | (uintptr_t)sq) & 15) != 0) return zc_fail(ctx, ZC_ERR_SIZE, "packed arrays must be 16-byte aligned");
  fe_mul_square_packed_kernel<ModP><<<grid_for(n, TPB), TPB, 0, ctx->stream>>>(a, b, prod, sq, n);
  ctx->launches++;
  ZC_CUDA(ctx, cudaGetLastError());
  return ZC_OK;
}
int32_t zc_fe_mul_square_batch_packed(zc_ctx* ctx, const uint8_t* a, const uint8_t* b, uint8_t* prod, uint8_t* sq, size_t n) {
  ZC_CHECK_CTX(ctx); ZC_CHECK_N(ctx, n);
  if (n == 0) return ZC_OK;
  ZC_CHECK_PTR(ctx, a && b && prod && sq);
  HostArr ins[2] = {{a, nullptr, 32, 0, nullptr}, {b, nullptr, 32, 1, nullptr}};
  HostArr outs[2] = {{nullptr, prod, 32, 2, nullptr}, {nullptr, sq, 32, 3, nullptr}};
  return host_pipelined(ctx, n, ins, 2, outs, 2, [&](size_t i0, size_t cnt) {
    return zc_fe_mul_square_batch_packed_dev(ctx, (const uint8_t*)ins[0].d + 32 * i0, (const uint8_t*)ins[1].d + 32 * i0,
                                             (uint8_t*)outs[0].d + 32 * i0, (uint8_t*)outs[1].d + 32 * i0, cnt);
  });
}

// ---- points ----------------------------------------------------------------------------------------------
#define ZC_DEFINE_PT_BIN(NAME, OP)                                                                                \
  int32_t NAME##_dev(zc_ctx* ctx, const uint64_t* p, const uint64_t* q, uint64_t* out, size_t n) {                \
    ZC_CHECK_CTX(ctx); ZC_CHECK_N(ctx, n);                                                                        \
    if (n == 0) return ZC_OK;                                                                                     \
    ZC_CHECK_PTR(ctx, p && q && out);                                                                             \
    int32_t rc_ = launch_pt<OP>(ctx, p, q, out, n);                                                               \
    return (rc_ == ZC_OK && ctx->validate) ? validation_finish(ctx, 4, nullptr) : rc_;                            \
  }                                                                                                               \
  int32_t NAME(zc_ctx* ctx, const uint64_t* p, const uint64_t* q, uint64_t* out, size_t n) {                      \
    ZC_CHECK_CTX(ctx); ZC_CHECK_N(ctx, n);                                                                        \
    if (n == 0) return ZC_OK;                                                                                     \
    ZC_CHECK_PTR(ctx, p && q && out);                                                                             \
    int32_t rc_ = host_binary(ctx, n, p, 160, q, 160, out, 160, [&](const void* da, const void* db, void* dout, size_t m) { \
      return launch_pt<OP>(ctx, (const uint64_t*)da, (const uint64_t*)db, (uint64_t*)dout, m);                    \
    });                                                                                                           \
    return (rc_ == ZC_OK && ctx->validate) ? validation_finish(ctx, 4, nullptr) : rc_;                            \
  }
#define ZC_DEFINE_PT_UN(NAME, OP)                                                                                 \
  int32_t NAME##_dev(zc_ctx* ctx, const uint64_t* p, uint64_t* out, size_t n) {                                   \
    ZC_CHECK_CTX(ctx); ZC_CHECK_N(ctx, n);                                                                        \
    if (n == 0) return ZC_OK;                                                                                     \
    ZC_CHECK_PTR(ctx, p && out);                                                                                  \
    int32_t rc_ = launch_pt<OP>(ctx, p, nullptr, out, n);                                                         \
    return (rc_ == ZC_OK && ctx->validate) ? validation_finish(ctx, 4, nullptr) : rc_;                            \
  }                                                                                                               \
  int32_t NAME(zc_ctx* ctx, const uint64_t* p, uint64_t* out, size_t n) {                                         \
    ZC_CHECK_CTX(ctx); ZC_CHECK_N(ctx, n);                                                                        \
    if (n == 0) return ZC_OK;                                                                                     \
    ZC_CHECK_PTR(ctx, p && out);                                                                                  \
    int32_t rc_ = host_binary(ctx, n, p, 160, nullptr, 0, out, 160, [&](const void* da, const void*, void* dout, size_t m) { \
      return launch_pt<OP>(ctx, (const uint64_t*)da, nullptr, (uint64_t*)dout, m);                                \
    });                                                                                                           \
    return (rc_ == ZC_OK && ctx->validate) ? validation_finish(ctx, 4, nullptr) : rc_;                            \
  }

ZC_DEFINE_PT_BIN(zc_point_add_batch, PT_ADD)
ZC_DEFINE_PT_BIN(zc_point_sub_batch, PT_SUB)
ZC_DEFINE_PT_UN(zc_point_double_batch, PT_DOUBLE)
ZC_DEFINE_PT_UN(zc_point_neg_batch, PT_NEG)

int32_t zc_ristretto_eq_batch_dev(zc_ctx* ctx, const uint64_t* p, const uint64_t* q, uint8_t* eq, size_t n) {
  ZC_CHECK_CTX(ctx); ZC_CHECK_N(ctx, n);
  if (n == 0) return ZC_OK;
  ZC_CHECK_PTR(ctx, p && q && eq);
  ristretto_eq_kernel<<<grid_for(n, TPB), TPB, 0, ctx->stream>>>(p, q, eq, n);
  ctx->launches++;
  ZC_CUDA(ctx, cudaGetLastError());
  return ZC_OK;
}
int32_t zc_ristretto_eq_batch(zc_ctx* ctx, const uint64_t* p, const uint64_t* q, uint8_t* eq, size_t n) {
  ZC_CHECK_CTX(ctx); ZC_CHECK_N(ctx, n);
  if (n == 0) return ZC_OK;
  ZC_CHECK_PTR(ctx, p && q && eq);
  return host_binary(ctx, n, p, 160, q, 160, eq, 1, [&](const void* da, const void* db, void* dout, size_t m) {
    return zc_ristretto_eq_batch_dev(ctx, (const uint64_t*)da, (const uint64_t*)db, (uint8_t*)dout, m);
  });
}

int32_t zc_point_scalar_mul_batch_dev(zc_ctx* ctx, const uint64_t* points, const uint64_t* scalars, uint64_t* out, size_t n, int32_t mode) {
  ZC_CHECK_CTX(ctx); ZC_CHECK_N(ctx, n);
  if (mode != ZC_SCALAR_MUL_STRICT && mode != ZC_SCALAR_MUL_FAST) return zc_fail(ctx, ZC_ERR_MODE, "unknown scalar-mul mode");
  if (n == 0) return ZC_OK;
  ZC_CHECK_PTR(ctx, points && scalars && out);
  if (ctx->validate) {
    int32_t rc;
    if ((rc = validation_enqueue<ModP>(ctx, points, 4 * n, 4 * ctx->vbase))) return rc;
    if ((rc = validation_enqueue<ModL>(ctx, scalars, n, ctx->vbase))) return rc;
  }
  if (mode == ZC_SCALAR_MUL_STRICT) {
    scalar_mul_strict_kernel<<<grid_for(n, STRICT_TPB), STRICT_TPB, STRICT_RING * 8 * STRICT_TPB * 16, ctx->stream>>>(points, scalars, out, n);
  } else {
    // persistent grid: 4 blocks per SM; table arena = one 1 KB table per resident thread
    unsigned grid = grid_for(n, SM_FAST_TPB);
    const unsigned max_grid = 4u * (unsigned)ctx->sm_count;
    if (grid > max_grid) grid = max_grid;
    const size_t nslots = (size_t)grid * SM_FAST_TPB;
    void* table = nullptr;
    int32_t rc = zc_scratch(ctx, 5, nslots * 1024, &table);
    if (rc) return rc;
    static const bool smem_table = getenv("ZC_SMF_SMEM") && atoi(getenv("ZC_SMF_SMEM")) != 0;
    if (smem_table) {
      static bool attr_set = false;
      if (!attr_set) { ZC_CUDA(ctx, cudaFuncSetAttribute(scalar_mul_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_FAST_TPB * 1024)); attr_set = true; }
      const unsigned g1 = grid < (unsigned)ctx->sm_count ? grid : (unsigned)ctx->sm_count;
      scalar_mul_fast_kernel<true><<<g1, SM_FAST_TPB, SM_FAST_TPB * 1024, ctx->stream>>>(points, scalars, out, n, nullptr, (size_t)g1 * SM_FAST_TPB);
    } else
    scalar_mul_fast_kernel<false><<<grid, SM_FAST_TPB, 0, ctx->stream>>>(points, scalars, out, n, (uint4*)table, nslots);
  }
  ctx->launches++;
  ZC_CUDA(ctx, cudaGetLastError());
  return (ctx->validate && !ctx->in_pipeline) ? validation_finish(ctx, 1, nullptr) : ZC_OK;
}
int32_t zc_point_scalar_mul_batch(zc_ctx* ctx, const uint64_t* points, const uint64_t* scalars, uint64_t* out, size_t n, int32_t mode) {
  ZC_CHECK_CTX(ctx); ZC_CHECK_N(ctx, n);
  if (mode != ZC_SCALAR_MUL_STRICT && mode != ZC_SCALAR_MUL_FAST) return zc_fail(ctx, ZC_ERR_MODE, "unknown scalar-mul mode");
  if (n == 0) return ZC_OK;
  ZC_CHECK_PTR(ctx, points && scalars && out);
  int32_t rc_ = host_binary(ctx, n, points, 160, scalars, 40, out, 160, [&](const void* da, const void* db, void* dout, size_t m) {
    return zc_point_scalar_mul_batch_dev(ctx, (const uint64_t*)da, (const uint64_t*)db, (uint64_t*)dout, m, mode);
  });
  return (rc_ == ZC_OK && ctx->validate) ? validation_finish(ctx, 1, nullptr) : rc_;
}

int32_t zc_point_fold_dev(zc_ctx* ctx, const uint64_t* points_dev, size_t k, uint64_t* out_point_dev) {
  ZC_CHECK_CTX(ctx);
  ZC_CHECK_PTR(ctx, out_point_dev && (k == 0 || points_dev));
  point_fold_kernel<<<1, 32, 0, ctx->stream>>>(points_dev, k, out_point_dev);
  ctx->launches++;
  ZC_CUDA(ctx, cudaGetLastError());
  return ZC_OK;
}


// ---- validation API ---------------------------------------------------------------------------------------------
int32_t zc_ctx_set_validation(zc_ctx* ctx, int32_t on) {
  ZC_CHECK_CTX(ctx);
  if (on) { int32_t rc = validation_setup(ctx); if (rc) return rc; }
  ctx->validate = on != 0;
  return ZC_OK;
}

#define ZC_DEFINE_CHECK(NAME, MOD, STRIDE_LIMBS)                                                                     \
  int32_t NAME##_dev(zc_ctx* ctx, const uint64_t* a, size_t n, uint64_t* first_bad) {                               \
    ZC_CHECK_CTX(ctx); ZC_CHECK_N(ctx, n);                                                                        \
    if (n == 0) return ZC_OK;                                                                                     \
    ZC_CHECK_PTR(ctx, a);                                                                                         \
    int32_t rc_ = validation_enqueue<MOD>(ctx, a, n * (STRIDE_LIMBS / 5), 0);                                     \
    if (rc_) return rc_;                                                                                          \
    unsigned long long fb = 0;                                                                                    \
    rc_ = validation_finish(ctx, STRIDE_LIMBS / 5, &fb);                                                          \
    if (rc_ == ZC_ERR_NONCANONICAL && first_bad) *first_bad = fb;                                                 \
    return rc_;                                                                                                   \
  }                                                                                                               \
  int32_t NAME(zc_ctx* ctx, const uint64_t* a, size_t n, uint64_t* first_bad) {                                   \
    ZC_CHECK_CTX(ctx); ZC_CHECK_N(ctx, n);                                                                        \
    if (n == 0) return ZC_OK;                                                                                     \
    ZC_CHECK_PTR(ctx, a);                                                                                         \
    void* d = nullptr;                                                                                            \
    int32_t rc_ = zc_scratch(ctx, 0, n * STRIDE_LIMBS * 8, &d);                                                   \
    if (rc_) return rc_;                                                                                          \
    ZC_CUDA(ctx, cudaMemcpyAsync(d, a, n * STRIDE_LIMBS * 8, cudaMemcpyHostToDevice, ctx->stream));               \
    return NAME##_dev(ctx, (const uint64_t*)d, n, first_bad);                                                     \
  }
ZC_DEFINE_CHECK(zc_fe_check_canonical_batch, ModP, 5)
ZC_DEFINE_CHECK(zc_scalar_check_canonical_batch, ModL, 5)
ZC_DEFINE_CHECK(zc_point_check_canonical_batch, ModP, 20)

}  // extern "C"

// internal (zc_msm.cu), C++ linkage: enqueue a canonical-input check (kind 1 field, 2 scalar, 3 point) / read the verdict
int32_t zc_validate_dev(zc_ctx* ctx, int32_t kind, const uint64_t* a, size_t n, size_t base) {   // internal (zc_msm.cu): 1 field, 2 scalar, 3 point
  if (kind == 2) return validation_enqueue<ModL>(ctx, a, n, base);
  return validation_enqueue<ModP>(ctx, a, kind == 3 ? 4 * n : n, kind == 3 ? 4 * base : base);
}
int32_t zc_validate_finish(zc_ctx* ctx, int32_t elems_per_unit) { return validation_finish(ctx, elems_per_unit, nullptr); }
