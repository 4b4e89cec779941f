// zc_msm.cu -- Pippenger multi-scalar multiplication  sum_i [s_i] P_i  on one GPU, or one rank's share of the
// windows when the MSM is sharded by bucket-window over several GPUs.
//
// The reference has no MSM: the semantics are fold(Add, identity, [double_and_add(P_i, s_i)])
// (/root/reference/src/edwards.rs:102-120, 465-489) as a group element.  Fast formulas are used throughout (cached-operand
// addition, dedicated doubling), so the result is compared canonically (affine / Ristretto equality), not limb-wise.
//
// Pipeline (recorded once into a CUDA graph per argument set; st = the context's stream):
//   prep     [side stream] points (AoS radix-2^52, normal form) -> cached operands (Y+X, Y-X, Z, 2dT) as 32 x u32.  A
//            normal-form coordinate vector IS a Montgomery-form representation of the same projective point (all four
//            coordinates scaled by 1/R), so no conversion multiply is needed -- only the 2d*T product.  Skipped for
//            prepared points (zc_msm_prepare_points_dev).
//   digits   [st] scalars -> signed c-bit digits for this rank's windows (+ per-bucket histogram, global atomics)
//   scan     [st] exclusive scan of the histogram per window (bucket start offsets)
//   scatter  [st] counting-sort scatter of (point index | sign) into bucket order (offset + the rank the histogram atomic returned)
//   then, per group of windows, top-down:
//   accum    [st] one thread per segment of the sorted list (balanced), 8M cached additions, cp.async-staged operands
//   fixq / heavy / cube1 / cube2a / cube2b   [one side stream per group] stitch buckets that span segments and reduce
//            sum_k (k+1) B_k by digit marginals -> four components per window
//   chain    [chain stream] fold the components into the running sum with Horner doublings, four lanes per operation
//   exchange (sharded only) partial points over NVLink peer memory fused with the fold (zc_peer.cu), or
//            ncclAllGather + fold kernel
#include <stdlib.h>
#include <vector>

#include "zc_internal.h"
#include "zc_point.cuh"
#include "zc_quad.cuh"

using namespace zc;

int32_t zc_nccl_allgather(zc_ctx *ctx, const void *send, void *recv, size_t bytes);   // zc_nccl.cu

namespace {

constexpr int MAX_WINDOWS = 32;      // ceil(256 / 8)
constexpr int MAX_GROUPS = 4;        // window groups processed top-down; the scaling chain of one group overlaps the next

// 1/d * R mod p: recovers 2T from the cached 2dT when a bucket is initialised from its first point
__device__ __forceinline__ Fe DINV_MONT() {
  return Fe{{0x69c50bb0u, 0xa53327e2u, 0x96b47422u, 0xeaa0ffd5u, 0xfd35fb8fu, 0xd34f1e03u, 0x8d35344bu, 0x0b7245f4u}};
}
// ---- prep ---------------------------------------------------------------------------------------------------
// points (ABI layout, 160 B) -> cached operands (Y+X, Y-X, Z, 2dT), 128 B.  One warp converts 32 consecutive points: the
// 5120 input bytes are read as 320 coalesced 16-byte loads into a per-warp shared-memory tile (176-byte point stride, so a
// lane's 8-byte limb reads spread over the banks), and the 4096 output bytes leave through the same tile as coalesced
// 16-byte stores.  The round-1 kernel read its point with twenty 8-byte loads at 160-byte stride and wrote 16-byte pieces
// at 128-byte stride (32 sectors per request either way): 97 us for 2^20 points against ~55 us of HBM time.
__global__ void __launch_bounds__(256) msm_prep_simple_kernel(const uint64_t* __restrict__ points, uint32_t* __restrict__ cached, size_t n) {
  size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;              // fallback for a point array that is not 16-byte aligned
  if (i >= n) return;
  PtCached c = pt_to_cached(pt_load52(points + 20 * i));
  uint32_t* o = cached + 32 * i;
  st_fe(o, c.YpX); st_fe(o + 8, c.YmX); st_fe(o + 16, c.Z); st_fe(o + 24, c.T2d);
}
constexpr int PREP_TPB = 128, PREP_STRIDE16 = 11;              // tile: 32 points x 11 uint4 per warp
__global__ void __launch_bounds__(PREP_TPB) msm_prep_kernel(const uint64_t* __restrict__ points, uint32_t* __restrict__ cached, size_t n) {
  __shared__ __align__(16) uint4 tile[PREP_TPB / 32][32 * PREP_STRIDE16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint4* t = tile[warp];
  // persistent: a few blocks per SM walk the tiles, so the pass never fills an SM and the (high-priority) sort kernels that run
  // beside it always find room -- with one block per 128 points the SMs' registers stayed full of this kernel's blocks and
  // every 512-thread sort block waited for the pass to END
  for (size_t base = ((size_t)blockIdx.x * (PREP_TPB / 32) + warp) * 32; base < n; base += (size_t)gridDim.x * PREP_TPB) {
  const size_t cnt = n - base < 32 ? n - base : 32;
  __syncwarp();
  const uint4* src = reinterpret_cast<const uint4*>(points + 20 * base);
#pragma unroll
  for (int k = 0; k < 10; k++) {
    const int j = lane + 32 * k;                                                   // 16-byte piece j of the warp's 320
    if ((size_t)j < cnt * 10) t[(j / 10) * PREP_STRIDE16 + (j % 10)] = __ldg(src + j);
  }
  __syncwarp();
  PtCached c;
  if ((size_t)lane < cnt) c = pt_to_cached(pt_load52(reinterpret_cast<const uint64_t*>(t + lane * PREP_STRIDE16)));
  __syncwarp();
  if ((size_t)lane < cnt) {
    uint4* o = t + lane * 8;                                                       // output tile: 128-byte stride
    o[0] = make_uint4(c.YpX.w[0], c.YpX.w[1], c.YpX.w[2], c.YpX.w[3]); o[1] = make_uint4(c.YpX.w[4], c.YpX.w[5], c.YpX.w[6], c.YpX.w[7]);
    o[2] = make_uint4(c.YmX.w[0], c.YmX.w[1], c.YmX.w[2], c.YmX.w[3]); o[3] = make_uint4(c.YmX.w[4], c.YmX.w[5], c.YmX.w[6], c.YmX.w[7]);
    o[4] = make_uint4(c.Z.w[0], c.Z.w[1], c.Z.w[2], c.Z.w[3]);         o[5] = make_uint4(c.Z.w[4], c.Z.w[5], c.Z.w[6], c.Z.w[7]);
    o[6] = make_uint4(c.T2d.w[0], c.T2d.w[1], c.T2d.w[2], c.T2d.w[3]); o[7] = make_uint4(c.T2d.w[4], c.T2d.w[5], c.T2d.w[6], c.T2d.w[7]);
  }
  __syncwarp();
  uint4* dst = reinterpret_cast<uint4*>(cached + 32 * base);
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int j = lane + 32 * k;
    if ((size_t)j < cnt * 8) dst[j] = t[j];
  }
  }
}

// prepared points: normalise to Z = 1 first (one inversion per point, paid once per point set), so that every later
// accumulation step is a 7-multiplication mixed addition.  Cached form (y+x, y-x, 1, 2dxy), all in Montgomery form.
__device__ __noinline__ Fe prep_mul(Fe a, Fe b) { return mont_mul<ModP>(a, b); }
// z^(p-2), Montgomery form (one chain per THREAD: the per-point inverses come from Montgomery's trick below)
__device__ __forceinline__ Fe prep_invert(const Fe& z) {
  const uint32_t e[8] = {0x5cf5d3ebu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0u, 0u, 0u, 0x10000000u};   // p - 2
  Fe zi = z;
#pragma unroll 1
  for (int bit = 251; bit >= 0; bit--) {
    zi = mont_sqr<ModP>(zi);
    if ((e[bit >> 5] >> (bit & 31)) & 1u) zi = prep_mul(zi, z);
  }
  return zi;
}
// affine cached form (y+x, y-x, 1, 2dxy) of a point given X, Y and 1/Z (all Montgomery form)
__device__ __forceinline__ void prep_store_affine(uint32_t* __restrict__ o, const Fe& X, const Fe& Y, const Fe& zi) {
  typedef ModP M;
  const Fe x = prep_mul(X, zi), y = prep_mul(Y, zi);
  st_fe(o, fe_add<M>(y, x)); st_fe(o + 8, fe_sub<M>(y, x)); st_fe(o + 16, Consts<M>::R1());
  st_fe(o + 24, prep_mul(prep_mul(x, y), D2_MONT()));
}
// A thread owns PREP_K points (index t + j T, coalesced): the Z's are multiplied together, the product is inverted ONCE
// (~320 products) and the individual inverses are peeled off (3 products per point) -- one inversion per point before.
// The loaded normal-form Z is used as it is (the Montgomery form of Z / R): the peeled value is R^2 / Z, one product by 1 on
// the running inverse (once per thread) turns it into R / Z, i.e. 1 / Z in Montgomery form for the X R, Y R of to_mont.
// Z = 0 (not a point) is treated as 1 so that it cannot poison its neighbours; its own entry is then garbage, as before.
constexpr int PREP_K = 8;
__global__ void __launch_bounds__(128) msm_prep_affine_kernel(const uint64_t* __restrict__ points, uint32_t* __restrict__ cached, size_t n) {
  typedef ModP M;
  const size_t T = (size_t)gridDim.x * 128, t = (size_t)blockIdx.x * 128 + threadIdx.x;
  Fe pref[PREP_K];
  Fe acc = Consts<M>::R1();
#pragma unroll
  for (int j = 0; j < PREP_K; j++) {
    const size_t i = t + (size_t)j * T;
    pref[j] = acc;
    if (i < n) {
      const Fe z = fe_load52(points + 20 * i + 10);
      if (!fe_is_zero(z)) acc = prep_mul(acc, z);
    }
  }
  Fe inv = prep_mul(prep_invert(acc), Fe{{1, 0, 0, 0, 0, 0, 0, 0}});
#pragma unroll
  for (int j = PREP_K - 1; j >= 0; j--) {
    const size_t i = t + (size_t)j * T;
    if (i < n) {
      const Fe z = fe_load52(points + 20 * i + 10);
      const Fe zi = prep_mul(inv, pref[j]);                          // R / Z_i: Montgomery form of 1 / Z_i
      if (!fe_is_zero(z)) inv = prep_mul(inv, z);
      prep_store_affine(cached + 32 * i, to_mont<M>(fe_load52(points + 20 * i)), to_mont<M>(fe_load52(points + 20 * i + 5)), zi);
    }
  }
}

// ---- digits + histogram ----------------------------------------------------------------------------------------
// digit d_w in [-2^(c-1), 2^(c-1)):  s = sum_w d_w 2^(c w).  Bucket slot = |d| - 1 in [0, 2^(c-1)).
// Recoding without a carry chain: with H = sum_w 2^(c-1) 2^(c w),  d_w = ((s + H) >> c w) mod 2^c  -  2^(c-1)
// (identical digits to the carry-propagating recode).  C is a template parameter so every shift is a constant and the
// scalar stays in registers.
// Canonical scalars are < L < 2^250, so a window that starts at bit c w >= 250 - (c-1) only ever sees digits in
// [0, 2^(250 - c w)]: a handful of buckets would take all n points (2^20 atomics on 2^9 addresses, thousands of
// additions per bucket).  Such a window spreads every digit over 2^SUB sub-buckets chosen by the low bits of the point
// index, so its histogram, scatter and accumulation look like any other window's; msm_fold_kernel sums the sub-buckets
// back before the reduction.
__host__ __device__ constexpr int short_window_sub_bits(int c, int w) {
  const int ba = 250 - c * w < 0 ? 0 : 250 - c * w;
  return ba < c - 1 ? (c - 1) - ba : 0;
}

// Fixed-base (merged) mode has one bucket set for all windows, so a short window cannot get sub-buckets of its own.  It is
// spread differently: its table rows are scaled by 2^(c w - SM) instead of 2^(c w) and its (non-negative) digit becomes
// d' = d 2^SM + (i mod 2^SM), which lands anywhere in the bucket range.  The extra  sum_i (i mod 2^SM) 2^(c w - SM) P_i  does
// not depend on the scalars: it is computed once with the tables and subtracted by the chain kernel.  SM = SUB - 1 keeps
// d' <= 2^(c-2) + 2^SM inside the bucket range for every digit a canonical scalar can produce.
__host__ __device__ constexpr int merged_spread_bits(int c, int w) {
  return (250 - c * w > 0 && short_window_sub_bits(c, w) > 1) ? short_window_sub_bits(c, w) - 1 : 0;
}

// Window ownership of the bucket-window sharding: boustrophedon over the ranks (windows 0..R-1 go to ranks 0..R-1, windows
// R..2R-1 to ranks R-1..0, and so on).  A rank's partial sum costs  reduce(top window) + c (w_top - w_low) doublings  while its
// lower windows accumulate, then  reduce(lowest window) + c w_low doublings  after the last accumulation: pairing the
// highest window with the lowest one keeps both serial chains short on every rank (w mod R put windows 7 and 15 on one
// rank at R = 8: 128 + 112 doublings; now (7, 8): 16 + 112, and (0, 15): 240 + 0 with the long chain hidden under the
// low window's accumulation).
__host__ __device__ constexpr int window_owner(int w, int nranks) {
  return ((w / nranks) & 1) ? nranks - 1 - (w % nranks) : (w % nranks);
}

// Which of this rank's tasks (local index tl, or -1) takes window w, and for which point range [p0, p1).  A task is a
// window (all points) today; the range exists so that a window can be split by points between ranks.
// merged != 0 (fixed-base tables): all of the rank's windows share ONE set of buckets -- the table entry of (window, point)
// is already scaled by 2^(c w), so only the digit's magnitude picks the bucket.
struct WinMap { int16_t tl[MAX_WINDOWS]; uint32_t p0[MAX_WINDOWS], p1[MAX_WINDOWS]; int32_t merged; };

// fixed-base tables: row (t, i) = 2^(c w_t) P_i in the same affine cached form, for the windows w_0 < w_1 < ... this rank
// owns.  One thread per point walks up the windows with dedicated doublings and normalises at every owned window (one
// inversion each): a one-time cost per generator set, after which an MSM needs neither per-window buckets nor any doubling.
struct FbWindows { int nwl; int16_t w[MAX_WINDOWS]; };      // a spread short window's rows are 2^(c w - SM) P_i (merged_spread_bits)
__device__ __noinline__ Pt prep_double(Pt p) { return pt_double_fast(p); }
__global__ void __launch_bounds__(128) msm_fixed_base_table_kernel(const uint64_t* __restrict__ points, uint32_t* __restrict__ table,
                                                                   size_t n, int c, const FbWindows win) {
  typedef ModP M;
  size_t i = (size_t)blockIdx.x * 128 + threadIdx.x;
  if (i >= n) return;
  Pt p = pt_to_mont(pt_load52(points + 20 * i));
  // Pass 1: walk up the windows; at every owned window park the UNNORMALISED (X, Y, Z) in the row itself, with the product of
  // the Z's before it in the row's fourth slot.  Pass 2: one inversion of the whole product, then Montgomery's trick backwards
  // over the rows -- 7 products per row + one chain per point instead of a chain per row (16 rows per point on one GPU).
  int at = 0;                                    // p = 2^at P_i
  Fe acc = Consts<M>::R1();
#pragma unroll 1
  for (int t = 0; t < win.nwl; t++) {
    const int target = c * (int)win.w[t] - merged_spread_bits(c, (int)win.w[t]);
#pragma unroll 1
    for (int j = at; j < target; j++) p = prep_double(p);
    at = target;
    uint32_t* o = table + 32 * ((size_t)t * n + i);
    st_fe(o, p.X); st_fe(o + 8, p.Y); st_fe(o + 16, p.Z); st_fe(o + 24, acc);
    if (!fe_is_zero(p.Z)) acc = prep_mul(acc, p.Z);
  }
  Fe inv = prep_invert(acc);
#pragma unroll 1
  for (int t = win.nwl - 1; t >= 0; t--) {
    uint32_t* o = table + 32 * ((size_t)t * n + i);
    Fe X, Y, Z, pre;
    ld_fe(o, X); ld_fe(o + 8, Y); ld_fe(o + 16, Z); ld_fe(o + 24, pre);
    const Fe zi = prep_mul(inv, pre);
    if (!fe_is_zero(Z)) inv = prep_mul(inv, Z);
    prep_store_affine(o, X, Y, zi);
  }
}

// scalars of the spread correction  K = sum_i (i mod 2^sm) 2^shift P_i  (radix-2^52 limbs)
__global__ void __launch_bounds__(256) msm_spread_scalars_kernel(uint64_t* __restrict__ scalars, size_t n, int sm, int shift) {
  size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const uint64_t r = (uint64_t)i & ((1ull << sm) - 1ull);
  const int limb = shift / 52, bit = shift % 52;
  uint64_t out[5] = {0, 0, 0, 0, 0};
  out[limb] = (r << bit) & ((1ull << 52) - 1ull);
  if (limb + 1 < 5) out[limb + 1] = bit ? (r >> (52 - bit)) : 0;
#pragma unroll
  for (int k = 0; k < 5; k++) scalars[5 * i + k] = out[k];
}


// The entries one scalar contributes: f(wl, key) for every window of the map, key = sign * (bucket slot + 1), 0 = no entry.
template <int C, class F>
__device__ __forceinline__ void msm_scalar_entries(const uint64_t* __restrict__ sp, uint32_t i, const WinMap& map, F&& f) {
  constexpr int NWIN = (256 + C - 1) / C;
  constexpr uint32_t HALF = 1u << (C - 1), MASK = (1u << C) - 1u, NB = HALF;
  const uint64_t l0 = sp[0], l1 = sp[1], l2 = sp[2], l3 = sp[3], l4 = sp[4];
  // 5 x 52-bit limbs -> 64-bit words, plus H (compile-time constant), 320 bits
  uint64_t v[5];
  v[0] = l0 | (l1 << 52);
  v[1] = (l1 >> 12) | (l2 << 40);
  v[2] = (l2 >> 24) | (l3 << 28);
  v[3] = (l3 >> 36) | (l4 << 16);
  v[4] = 0;
  {
    uint64_t h[5] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int w = 0; w < NWIN; w++) { const int bit = C - 1 + C * w; h[bit >> 6] |= 1ull << (bit & 63); }
    uint64_t carry = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
      uint64_t t = v[k] + h[k];
      uint64_t c1 = t < v[k];
      v[k] = t + carry;
      carry = c1 | (v[k] < t);
    }
  }
#pragma unroll
  for (int w = 0; w < NWIN; w++) {
    const int wl = map.tl[w];
    if (wl >= 0) {
      const int bit = C * w, word = bit >> 6, sh = bit & 63;
      uint64_t x = v[word] >> sh;
      if (sh + C > 64 && word + 1 < 5) x |= v[word + 1] << (64 - sh);
      const int32_t d = (int32_t)((uint32_t)x & MASK) - (int32_t)HALF;
      int32_t key = 0;                                        // sign * (bucket slot + 1), 0 = skip
      if (d != 0 && i >= map.p0[w] && i < map.p1[w]) {
        uint32_t slot = (uint32_t)(d < 0 ? -d : d) - 1u;
        const int SUB = short_window_sub_bits(C, w);      // top window(s) of a 250-bit scalar: few distinct digits
        if (SUB > 0 && !map.merged) slot = ((slot << SUB) | ((uint32_t)i & ((1u << SUB) - 1u))) & (NB - 1u);
        key = d < 0 ? -(int32_t)(slot + 1u) : (int32_t)(slot + 1u);
      }
      const int SM = merged_spread_bits(C, w);
      if (SM > 0 && map.merged && i >= map.p0[w] && i < map.p1[w]) {
        const int32_t dd = d * (1 << SM) + (int32_t)((uint32_t)i & ((1u << SM) - 1u));
        key = 0;
        if (dd != 0) {
          const uint32_t slot = ((uint32_t)(dd < 0 ? -dd : dd) - 1u) & (NB - 1u);
          key = dd < 0 ? -(int32_t)(slot + 1u) : (int32_t)(slot + 1u);
        }
      }
      f(wl, key);
    }
  }
}

// Sort A (ZC_MSM_SORT=atomic; the round-1 sort): one thread per scalar, a global histogram atomic per entry which also
// hands out the entry's rank inside its bucket, then msm_scan_kernel and msm_scatter_kernel.
template <int C>
__global__ void __launch_bounds__(256) msm_digits_kernel(const uint64_t* __restrict__ scalars, size_t n, const WinMap map,
                                                         int32_t* __restrict__ digits, uint32_t* __restrict__ ranks,
                                                         uint32_t* __restrict__ hist) {
  constexpr uint32_t NB = 1u << (C - 1);
  size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  msm_scalar_entries<C>(scalars + 5 * i, (uint32_t)i, map, [&](int wl, int32_t key) {
    if (key != 0) {
      const uint32_t slot = (uint32_t)(key < 0 ? -key : key) - 1u;
      ranks[(size_t)wl * n + i] = atomicAdd(&hist[(map.merged ? (size_t)0 : (size_t)wl * NB) + slot], 1u);
    }
    digits[(size_t)wl * n + i] = key;
  });
}

// Sort B (default): a two-level counting sort with no global atomics.  The 2^(c-1) buckets of a bucket set ("problem": one
// local window, or the merged set of the fixed-base path) are cut into 2^CB coarse bins of 2^FB buckets.
//   count   every block walks its contiguous chunk of the scalars and counts its entries per (problem, bin) in shared memory
//   bscan   one warp per bin: exclusive scan of the per-block counts (where each block's entries start inside the bin) + bin total
//   place   the same walk again (the scalars are L2-resident; recomputing the digits is cheaper than storing them): an entry
//           goes to  bin start + block's offset + its rank among the block's entries of that bin (shared-memory atomic)
//   fine    one block per bin: histogram of the 2^FB buckets in shared memory, scan, write hist[] / offs[] for the bin's
//           buckets and the entries in bucket order
// Entry order inside a bucket is arbitrary (as with sort A); any digit distribution is handled (a bin that takes every
// entry is sorted by one block, slowly but correctly).
struct SortPlan { int cb, fb; uint32_t chunk; int lo; };      // coarse / fine bits, scalars per block, first local window of the sort
constexpr int SORT_TPB = 512;
constexpr int FINE_TPB = 512, FINE_CAP = 9216;                 // msm_sort_fine_kernel sorts a bin of up to FINE_CAP entries in shared memory (12 B each, 108 KiB: two blocks per SM)
// coarse bits: bins of ~2^13 entries (one msm_sort_fine block each, staged in shared memory), at most 2^8 buckets per bin
static int sort_coarse_bits(int bits, size_t entries_per_problem, int nprob_max) {
  int cb = 0;
  while (cb < 9 && ((size_t)8192 << cb) < entries_per_problem) cb++;
  if (cb < bits - 8) cb = bits - 8;
  if (cb > bits) cb = bits;
  while (cb > 0 && cb > bits - 8 && ((size_t)nprob_max << cb) > 4096) cb--;
  return cb;
}

// counts[bin][block]
template <int C>
__global__ void __launch_bounds__(SORT_TPB) msm_sort_count_kernel(const uint64_t* __restrict__ scalars, size_t n, const WinMap map,
                                                                  const SortPlan pl, int nbins, uint32_t* __restrict__ counts) {
  extern __shared__ uint32_t sort_sm[];
  for (int t = threadIdx.x; t < nbins; t += SORT_TPB) sort_sm[t] = 0;
  __syncthreads();
  const uint32_t beg = blockIdx.x * pl.chunk, end = (size_t)beg + pl.chunk < n ? beg + pl.chunk : (uint32_t)n;
  for (uint32_t i = beg + threadIdx.x; i < end; i += SORT_TPB)
    msm_scalar_entries<C>(scalars + 5 * (size_t)i, i, map, [&](int wl, int32_t key) {
      if (key == 0) return;
      const uint32_t slot = (uint32_t)(key < 0 ? -key : key) - 1u;
      const uint32_t prob = map.merged ? 0u : (uint32_t)(wl - pl.lo);
      atomicAdd(&sort_sm[(prob << pl.cb) | (slot >> pl.fb)], 1u);
    });
  __syncthreads();
  for (int t = threadIdx.x; t < nbins; t += SORT_TPB) counts[(size_t)t * gridDim.x + blockIdx.x] = sort_sm[t];
}

// one warp per bin: lane l owns the blocks [l per, (l+1) per) -- all loads in flight at once, one shuffle scan
constexpr int BSCAN_PER = 20;                                 // >= ceil(max blocks / 32): up to 640 blocks
__global__ void __launch_bounds__(128) msm_sort_bscan_kernel(uint32_t* __restrict__ counts, int nblk, int nbins, uint32_t* __restrict__ totals) {
  const int bin = (int)((blockIdx.x * 128 + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (bin >= nbins) return;
  uint32_t* row = counts + (size_t)bin * nblk;
  const int per = (nblk + 31) / 32, b0 = lane * per;
  uint32_t x[BSCAN_PER];
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < BSCAN_PER; k++) { x[k] = (k < per && b0 + k < nblk) ? row[b0 + k] : 0u; sum += x[k]; }
  uint32_t inc = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
  uint32_t run = inc - sum;
#pragma unroll
  for (int k = 0; k < BSCAN_PER; k++) if (k < per && b0 + k < nblk) { row[b0 + k] = run; run += x[k]; }
  if (lane == 31) totals[bin] = inc;
}

// tmp[pos] = fine key << 32 | entry
template <int C>
__global__ void __launch_bounds__(SORT_TPB) msm_sort_place_kernel(const uint64_t* __restrict__ scalars, size_t n, size_t n_pad, const WinMap map,
                                                                  const SortPlan pl, int nbins, const uint32_t* __restrict__ counts,
                                                                  const uint32_t* __restrict__ totals, uint32_t* __restrict__ binstart,
                                                                  uint64_t* __restrict__ tmp) {
  extern __shared__ uint32_t sort_sm[];
  uint32_t* base = sort_sm;                    // where this block's next entry of a bin goes (within the problem's array)
  uint32_t* cnt = sort_sm + nbins;             // the bin totals
  for (int t = threadIdx.x; t < nbins; t += SORT_TPB) cnt[t] = totals[t];
  __syncthreads();
  for (int t = threadIdx.x; t < nbins; t += SORT_TPB) {
    uint32_t s = 0;
    for (int u = t & ~((1 << pl.cb) - 1); u < t; u++) s += cnt[u];
    if (blockIdx.x == 0) binstart[t] = s;
    base[t] = s + counts[(size_t)t * gridDim.x + blockIdx.x];
  }
  __syncthreads();
  const uint32_t beg = blockIdx.x * pl.chunk, end = (size_t)beg + pl.chunk < n ? beg + pl.chunk : (uint32_t)n;
  for (uint32_t i = beg + threadIdx.x; i < end; i += SORT_TPB)
    msm_scalar_entries<C>(scalars + 5 * (size_t)i, i, map, [&](int wl, int32_t key) {
      if (key == 0) return;
      const uint32_t slot = (uint32_t)(key < 0 ? -key : key) - 1u;
      const uint32_t sign = key < 0 ? 0x80000000u : 0u;
      const uint32_t prob = map.merged ? 0u : (uint32_t)(wl - pl.lo);
      const uint32_t bin = (prob << pl.cb) | (slot >> pl.fb);
      const size_t pos = (size_t)prob * n_pad + atomicAdd(&base[bin], 1u);     // (a per-entry rank kept by the count pass instead of this atomic measured slower: 24.6 vs 19.1 us)
      // merged: one bucket set, the entry names the table row  wl * n + i
      const uint32_t e = (map.merged ? (uint32_t)((size_t)wl * n + i) : (uint32_t)i) | sign;
      tmp[pos] = ((uint64_t)(slot & ((1u << pl.fb) - 1u)) << 32) | e;
    });
}

__global__ void __launch_bounds__(FINE_TPB, 2) msm_sort_fine_kernel(const uint64_t* __restrict__ tmp, const uint32_t* __restrict__ binstart,
                                                                    const uint32_t* __restrict__ totals, const SortPlan pl, size_t n_pad, int nb,
                                                                    uint32_t* __restrict__ hist, uint32_t* __restrict__ offs, uint32_t* __restrict__ sorted) {
  extern __shared__ __align__(16) uint64_t fine_sm[];           // the bin's entries (FINE_CAP + 2), then the FINE_CAP sorted entries
  uint32_t* fine_out = reinterpret_cast<uint32_t*>(fine_sm + FINE_CAP + 2);
  __shared__ uint32_t h[256], cur[256], wsum[8];
  const int bin = blockIdx.x, prob = bin >> pl.cb, bl = bin & ((1 << pl.cb) - 1), nf = 1 << pl.fb;
  const uint32_t start = binstart[bin], cnt = totals[bin], jend = start + cnt, j0 = start & ~1u, skip = start - j0;
  const uint64_t* src = tmp + (size_t)prob * n_pad;             // 16-byte aligned (n_pad is a multiple of 32)
  const bool fits = jend - j0 <= (uint32_t)FINE_CAP;            // block-uniform
  if (threadIdx.x < 256) h[threadIdx.x] = 0;
  if (fits) {
#pragma unroll 4
    for (uint32_t j = j0 + 2u * threadIdx.x; j < jend; j += 2u * FINE_TPB) {
      if (j + 1 < jend) *reinterpret_cast<ulonglong2*>(fine_sm + (j - j0)) = *reinterpret_cast<const ulonglong2*>(src + j);
      else fine_sm[j - j0] = src[j];                              // the slot after the last bin's last entry was never written
    }
  }
  __syncthreads();
  if (fits) {                                    // the histogram atomic also ranks the entry inside its bucket: kept beside the key
    for (uint32_t j = threadIdx.x; j < cnt; j += FINE_TPB) {
      const uint64_t x = fine_sm[skip + j];
      const uint32_t key = (uint32_t)(x >> 32);
      fine_sm[skip + j] = (x & 0xffffffffull) | ((uint64_t)((atomicAdd(&h[key], 1u) << 8) | key) << 32);
    }
  }
  else      { for (uint32_t j = start + threadIdx.x; j < jend; j += FINE_TPB) atomicAdd(&h[(uint32_t)(src[j] >> 32)], 1u); }
  __syncthreads();
  if (threadIdx.x < 256) {                       // exclusive scan of h[0..255] (entries >= nf are zero)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t x = h[threadIdx.x];
    uint32_t inc = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) wsum[wid] = inc;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    uint32_t pre = 0;
    for (int u = 0; u < wid; u++) pre += wsum[u];
    const uint32_t ex = pre + inc - x;
    cur[threadIdx.x] = ex;
    if ((int)threadIdx.x < nf) {
      const size_t b = (size_t)prob * nb + ((size_t)bl << pl.fb) + threadIdx.x;
      hist[b] = x;
      offs[b] = start + ex;
    }
  }
  __syncthreads();
  uint32_t* out = sorted + (size_t)prob * n_pad + start;
  // the accumulation loads its segment's indices four at a time: the group that holds the problem's last entry reaches up to
  // three slots past it (never used) -- make them defined
  if (bl == (1 << pl.cb) - 1 && threadIdx.x < 4 && (size_t)start + cnt + threadIdx.x < n_pad) out[cnt + threadIdx.x] = 0u;   // stay inside this problem's array
  if (fits) {
    // bucket order is made in shared memory; the bin then leaves in one coalesced copy (scattered 4-byte stores cost a 32-byte
    // sector each at the L2: the kernel was bound by them)
    for (uint32_t j = threadIdx.x; j < cnt; j += FINE_TPB) { const uint64_t x = fine_sm[skip + j]; const uint32_t kr = (uint32_t)(x >> 32); fine_out[cur[kr & 255u] + (kr >> 8)] = (uint32_t)x; }
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < cnt; j += FINE_TPB) out[j] = fine_out[j];
  }
  else      { for (uint32_t j = start + threadIdx.x; j < jend; j += FINE_TPB) { const uint64_t x = src[j]; out[atomicAdd(&cur[(uint32_t)(x >> 32)], 1u)] = (uint32_t)x; } }
}

// ---- exclusive scan of each window's histogram (one block of SCAN_TPB threads per local window) -----------------------
// Thread t owns PER = nb / SCAN_TPB consecutive counters (16-byte loads), warps scan by shuffle, one shared-memory hop.
constexpr int SCAN_TPB = 1024;
__global__ void __launch_bounds__(SCAN_TPB) msm_scan_kernel(const uint32_t* __restrict__ hist, uint32_t* __restrict__ offs, int nb) {
  __shared__ uint32_t wtot[32];
  const int wl = blockIdx.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const uint32_t* h = hist + (size_t)wl * nb;
  const int per = nb >= 4 * SCAN_TPB ? nb / SCAN_TPB : 4;      // multiple of 4 (nb is a power of two >= 128)
  const int lo = t * per;
  uint32_t sum = 0;
  if (lo < nb) for (int k = 0; k < per; k += 4) { uint4 q = *reinterpret_cast<const uint4*>(h + lo + k); sum += q.x + q.y + q.z + q.w; }
  uint32_t inc = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
  if (lane == 31) wtot[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    uint32_t x = wtot[lane], y = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, y, d); if (lane >= d) y += o; }
    wtot[lane] = y - x;                                         // exclusive warp offsets
  }
  __syncthreads();
  uint32_t run = wtot[wid] + inc - sum;
  if (lo < nb) for (int k = 0; k < per; k += 4) {
    uint4 q = *reinterpret_cast<const uint4*>(h + lo + k);
    uint4 o; o.x = run; o.y = run + q.x; o.z = o.y + q.y; o.w = o.z + q.z; run = o.w + q.w;
    *reinterpret_cast<uint4*>(offs + (size_t)wl * nb + lo + k) = o;
  }
}

// ---- scatter: counting sort of point indices into bucket order ----------------------------------------------------
__global__ void __launch_bounds__(256) msm_scatter_kernel(const int32_t* __restrict__ digits, const uint32_t* __restrict__ ranks,
                                                          size_t n, size_t n_pad, int nwl, int nb, int merged,
                                                          const uint32_t* __restrict__ offs, uint32_t* __restrict__ sorted) {
  size_t g = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (g >= n * (size_t)nwl) return;
  size_t wl = g / n, i = g - wl * n;
  int32_t d = digits[g];
  if (d == 0) return;
  uint32_t slot = (uint32_t)(d < 0 ? -d : d) - 1u;
  const uint32_t sign = d < 0 ? 0x80000000u : 0u;
  if (merged) {                                  // one bucket set; the entry names the table row  wl * n + i
    sorted[offs[slot] + ranks[g]] = (uint32_t)g | sign;
    return;
  }
  const uint32_t pos = offs[wl * nb + slot] + ranks[g];
  sorted[wl * n_pad + pos] = (uint32_t)i | sign;
}

// ---- bucket accumulation, balanced: one thread per SEGMENT of SEG consecutive sorted entries ------------------------
// A thread walks its SEG entries, summing runs of equal bucket.  A run that covers its whole bucket is stored straight
// into buckets[]; a run cut by a segment boundary goes to the segment's H slot (run containing the segment's first
// entry) or T slot (run containing its last entry, when that is a different run).  load_bucket / msm_heavy_kernel then
// stitch the buckets that span several segments.  Every thread does the same number of additions, whatever the digit
// distribution (a top window with few, heavy buckets used to serialise thousands of additions in one thread).
constexpr int SEG_MAX = 32;           // segment length is 8, 16 or 32 (chosen per call from the amount of work)
// threads per accumulation block: 128 (ZC_MSM_ACC_TPB=64 selects 64-thread blocks, an experiment: a launch of 512 blocks on 592
// slots leaves the SMs unevenly loaded, but finer blocks measured the same time -- the kernel is bound by the dependent
// multiplications of each thread's segment, not by the busiest SM)
constexpr int ACC_TPB_MAX = 128;

// Operands are gathered into shared memory with cp.async (no register staging) one entry ahead of the addition that
// uses them, and the addition reads each 32-byte coordinate from shared memory right before the multiplication that
// consumes it: the kernel holds the accumulator and one product's temporaries in registers instead of the accumulator
// plus two whole operands (166 -> <= 128 registers, 3 -> 4 resident CTAs per SM).
// Stage layout: [buffer 0..1][16-byte piece 0..7][thread]; pieces 0-1 = Y+X, 2-3 = Y-X, 4-5 = Z, 6-7 = 2dT.
constexpr int ACC_NBUF = 2;

template <int ACC_TPB>
__device__ __forceinline__ void gather_entry(uint4* __restrict__ stage_buf, const uint32_t* __restrict__ cached, uint32_t e, int tx) {
  const uint32_t* src = cached + 32 * (size_t)(e & 0x7fffffffu);
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(stage_buf + tx);
#pragma unroll
  for (int k = 0; k < 8; k++)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)(k * ACC_TPB * 16)), "l"(src + 4 * k) : "memory");
}
template <int ACC_TPB>
__device__ __forceinline__ Fe lds_fe(const uint4* __restrict__ p) {      // two pieces, ACC_TPB apart
  const uint4 lo = p[0], hi = p[ACC_TPB];
  return Fe{{lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w}};
}
// One out-of-line copy of the multiplier: the accumulation loop body shrinks from ~35 KB of straight-line code to ~6 KB.
// With 16 warps per SM at different points of the loop, instruction fetch was the top stall reason (no_instruction 1.3
// per issue -> 0.07); alone the kernel is 4 % slower for the call overhead, inside the MSM (sharing SMs with the side
// stream, 104 registers instead of 125) the whole 2^20-point MSM is 3 % faster.  A rolled-loop inline multiplier
// (mont_mul_rolled) was slower on both counts.
__device__ __noinline__ Fe acc_mul(Fe a, Fe b) { return mont_mul<ModP>(a, b); }

// p + (+-q) with q staged in shared memory (q points at this thread's piece 0); add-2008-hwcd-3, a = -1
// AFFINE: the staged operands have Z = 1 (prepared points are normalised once), so D = 2 Z1 needs no product: 7M.
template <bool AFFINE, int ACC_TPB>
__device__ __forceinline__ Pt pt_add_staged(const Pt& p, const uint4* __restrict__ q, bool neg) {
  typedef ModP M;
  // p and the staged operand are canonical; every linear combination below is lazy (no conditional subtraction): the
  // factors stay below 2m, 2m, 3m, 3m and each product below 9 m^2 < R m, which is all the Montgomery product needs
  const Fe zero{{0, 0, 0, 0, 0, 0, 0, 0}};
  // ZC_ACC_INLINE: bit k set = product k of the addition (A, B, C, D, X3, Y3, Z3, T3) is inlined instead of going through
  // the out-of-line multiplier.  Every call costs ~16 IMAD.MOV of register marshalling on the multiplier pipe; every inlined
  // product ~3.7 KB of code.  Measured at 2^20 points (one GPU, prepared): 0x00 2.758 ms, 0xf0 (the four output products)
  // 2.630 ms; the fully inlined loop of round 1 (35 KB) ran out of the 32 KB instruction cache.
#ifndef ZC_ACC_INLINE
#define ZC_ACC_INLINE 0xf0
#endif
#define ZC_ACC_MUL(K, X, Y) (((ZC_ACC_INLINE >> (K)) & 1) ? mont_mul<ModP>((X), (Y)) : acc_mul((X), (Y)))
  Fe A = ZC_ACC_MUL(0, fe_sub_lazy<1>(p.Y, p.X), lds_fe<ACC_TPB>(q + (neg ? 0 : 2) * ACC_TPB));
  Fe B = ZC_ACC_MUL(1, fe_add_lazy(p.Y, p.X), lds_fe<ACC_TPB>(q + (neg ? 2 : 0) * ACC_TPB));
  Fe t2d = lds_fe<ACC_TPB>(q + 6 * ACC_TPB);
  if (neg) t2d = fe_sub_lazy<1>(zero, t2d);                    // m - 2dT2 in (0, m]
  Fe C = ZC_ACC_MUL(2, p.T, t2d);
  Fe D = AFFINE ? p.Z : ZC_ACC_MUL(3, p.Z, lds_fe<ACC_TPB>(q + 4 * ACC_TPB));
  D = fe_dbl_lazy(D);                                          // < 2m
  Fe E = fe_sub_lazy<1>(B, A);                                 // < 2m
  Fe F = fe_sub_lazy<1>(D, C);                                 // < 3m
  Fe G = fe_add_lazy(D, C);                                    // < 3m
  Fe H = fe_add_lazy(B, A);                                    // < 2m
  Pt r;
  r.X = ZC_ACC_MUL(4, E, F);
  r.Y = ZC_ACC_MUL(5, G, H);
  r.Z = ZC_ACC_MUL(6, F, G);
  r.T = ZC_ACC_MUL(7, E, H);
#undef ZC_ACC_MUL
  return r;
}
// the point a staged operand stands for, as (2X, 2Y, 2Z, 2T)
template <int ACC_TPB>
__device__ __forceinline__ Pt staged_to_pt(const uint4* __restrict__ q, bool neg) {
  typedef ModP M;
  Fe ypx = lds_fe<ACC_TPB>(q + (neg ? 2 : 0) * ACC_TPB), ymx = lds_fe<ACC_TPB>(q + (neg ? 0 : 2) * ACC_TPB);
  Fe z = lds_fe<ACC_TPB>(q + 4 * ACC_TPB), t2d = lds_fe<ACC_TPB>(q + 6 * ACC_TPB);
  if (neg) t2d = fe_neg<M>(t2d);
  Pt r;
  r.X = fe_sub<M>(ypx, ymx);
  r.Y = fe_add<M>(ypx, ymx);
  r.Z = fe_add<M>(z, z);
  r.T = acc_mul(t2d, DINV_MONT());
  return r;
}

template <bool AFFINE, int ACC_TPB>
__global__ void __launch_bounds__(ACC_TPB, 512 / ACC_TPB) msm_accum_kernel(const uint32_t* __restrict__ cached, const uint32_t* __restrict__ sorted,
                                                               const uint32_t* __restrict__ offs, const uint32_t* __restrict__ hist,
                                                               size_t n_pad, int nseg, int seg, int nwl, int nb,
                                                               uint32_t* __restrict__ buckets, uint32_t* __restrict__ partH,
                                                               uint32_t* __restrict__ partT, uint32_t g_lo, uint32_t g_hi) {
  __shared__ uint32_t idx_s[SEG_MAX * ACC_TPB];
  __shared__ __align__(16) uint4 stage[ACC_NBUF][8 * ACC_TPB];
  const int tx = threadIdx.x;
  size_t g = (size_t)g_lo + (size_t)blockIdx.x * ACC_TPB + tx;      // this launch covers the segments [g_lo, g_hi) of nwl * nseg
  if (g >= (size_t)g_hi) return;
  const size_t wl = g / nseg;
  const uint32_t s = (uint32_t)(g - wl * nseg);
  const uint32_t* woffs = offs + wl * nb;
  const uint32_t* whist = hist + wl * nb;
  const uint32_t nnz = woffs[nb - 1] + whist[nb - 1];
  const uint32_t start = s * (uint32_t)seg;
  if (start >= nnz) return;
  const uint32_t end = min(start + (uint32_t)seg, nnz);
  if ((seg & 3) == 0) {
    const uint4* src = reinterpret_cast<const uint4*>(sorted + wl * n_pad + start);
#pragma unroll 2
    for (int j = 0; j < seg / 4; j++) {
      if (start + 4u * j >= end) break;                         // nothing of this group is used
      uint4 v = src[j];
      idx_s[(4 * j + 0) * ACC_TPB + tx] = v.x; idx_s[(4 * j + 1) * ACC_TPB + tx] = v.y;
      idx_s[(4 * j + 2) * ACC_TPB + tx] = v.z; idx_s[(4 * j + 3) * ACC_TPB + tx] = v.w;
    }
  } else {                                   // even segment lengths that are not multiples of four (wave-filling lengths such as 14)
    const uint2* src = reinterpret_cast<const uint2*>(sorted + wl * n_pad + start);
#pragma unroll 2
    for (int j = 0; j < seg / 2; j++) {
      if (start + 2u * j >= end) break;
      uint2 v = src[j];
      idx_s[(2 * j + 0) * ACC_TPB + tx] = v.x; idx_s[(2 * j + 1) * ACC_TPB + tx] = v.y;
    }
  }
  uint32_t e_cur = idx_s[tx];
  gather_entry<ACC_TPB>(stage[0], cached, e_cur, tx);
  asm volatile("cp.async.commit_group;" ::: "memory");
  // bucket of the first entry: last b with offs[b] <= start  (upper_bound - 1)
  uint32_t lo = 0, hi = (uint32_t)nb;
  while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (woffs[mid] <= start) lo = mid + 1; else hi = mid; }
  uint32_t b = lo - 1;
  uint32_t bbeg = woffs[b], bend = bbeg + whist[b];
  uint32_t run_start = start;
  uint32_t* bk = buckets + 32 * (wl * nb);
  uint32_t* pH = partH + 32 * g;
  uint32_t* pT = partT + 32 * g;

  Pt acc;
#pragma unroll 1
  for (uint32_t k = start; k < end; k++) {
    const int buf = (k - start) & 1;
    uint32_t e_nxt = 0;
    if (k + 1 < end) {                       // gather the next operand under this addition
      e_nxt = idx_s[(k + 1 - start) * ACC_TPB + tx];
      gather_entry<ACC_TPB>(stage[buf ^ 1], cached, e_nxt, tx);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");     // everything but the newest group: entry k has landed
    const uint4* q = stage[buf] + tx;
    const bool neg = (e_cur >> 31) != 0;
    if (k == start) {
      acc = staged_to_pt<ACC_TPB>(q, neg);
    } else if (k == bend) {
      // flush the finished run
      const bool complete = (run_start == bbeg);
      st_pt(complete ? bk + 32 * (size_t)b : pH, acc);   // incomplete here <=> the run began before this segment (H slot)
      do { b++; } while (whist[b] == 0);
      bbeg = k; bend = k + whist[b];
      run_start = k;
      acc = staged_to_pt<ACC_TPB>(q, neg);
    } else {
      acc = pt_add_staged<AFFINE, ACC_TPB>(acc, q, neg);
    }
    e_cur = e_nxt;
  }
  {
    const bool complete = (run_start == bbeg) && (end == bend);
    uint32_t* dst = complete ? bk + 32 * (size_t)b : (run_start == start ? pH : pT);
    st_pt(dst, acc);
  }
}

// ---- buckets that span several segments ---------------------------------------------------------------------------
// bucket range [o, e): s_first = o / SEG, s_last = (e-1) / SEG.  If s_first == s_last the run was complete and is already
// in buckets[].  Otherwise  sum = (o == s_first*SEG ? H[s_first] : T[s_first]) + H[s_first+1] + ... + H[s_last].
// Buckets with at most FIX_INLINE partials are stitched on the fly by whoever reads them (load_bucket, used by the
// first reduction stage); heavier ones are queued here for msm_heavy_kernel (one warp per bucket, tree sum), which
// writes them into buckets[].
static const int FIX_INLINE = getenv("ZC_MSM_FIX_INLINE") ? atoi(getenv("ZC_MSM_FIX_INLINE")) : 8;       // per-window buckets; the merged buckets of the fixed-base path choose theirs per call
__global__ void __launch_bounds__(256) msm_fixq_kernel(const uint32_t* __restrict__ offs, const uint32_t* __restrict__ hist,
                                                       int seg, int nwl, int nb, int fix_inline, uint32_t* __restrict__ heavy_count, uint32_t* __restrict__ heavy_list,
                                                       long long lim_lo, long long lim_hi) {
  size_t g = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (g >= (size_t)nwl * nb) return;
  const uint32_t cnt = hist[g];
  if (cnt == 0) return;
  const uint32_t o = offs[g], e = o + cnt;
  if ((long long)e <= lim_lo || (long long)e > lim_hi) return;
  const uint32_t s_first = o / (uint32_t)seg, s_last = (e - 1) / (uint32_t)seg;
  if (s_last - s_first + 1 > (uint32_t)fix_inline) {
    uint32_t slot = atomicAdd(heavy_count, 1u);
    heavy_list[slot] = (uint32_t)g;
  }
}

// Out-of-line point operations for the latency-bound tail kernels (reduce / heavy): one copy of the ~30 KB addition
// body per kernel keeps them inside the instruction cache (inlined at every call site, msm_reduce1_kernel was 1 MB of
// straight-line code and ran at ~8 cycles per instruction).
__device__ __noinline__ Pt pt_add_ni(Pt p, Pt q) { return pt_add_fast(p, q); }
__device__ __noinline__ Pt pt_double_ni(Pt p) { return pt_double_fast(p); }

struct BucketSrc {                  // where a group's buckets live (all pointers relative to the group's first window)
  const uint32_t *offs, *hist, *partH, *partT, *buckets;
  int nseg, nb, seg, fix_inline;
  // split accumulation (one bucket set accumulated by several launches over consecutive ranges of the sorted list): this
  // launch owns the buckets whose last entry lies in (lim_lo, lim_hi] -- complete once the launch up to lim_hi has finished
  long long lim_lo, lim_hi;
};
constexpr long long LIM_ALL_LO = -1, LIM_ALL_HI = 0x7fffffffffffffffll;
// bucket g of the group, stitched from its segment partials when it spans a few segments
__device__ __forceinline__ Pt load_bucket(const BucketSrc& b, size_t g) {
  const int seg = b.seg;
  const uint32_t cnt = b.hist[g];
  if (cnt == 0) return pt_identity_mont();
  const uint32_t o = b.offs[g], e = o + cnt;
  const uint32_t s_first = o / (uint32_t)seg, s_last = (e - 1) / (uint32_t)seg;
  if (s_first == s_last || s_last - s_first + 1 > (uint32_t)b.fix_inline) return ld_pt(b.buckets + 32 * g);
  const size_t wl = g / b.nb;
  const uint32_t* H = b.partH + 32 * (wl * b.nseg);
  const uint32_t* T = b.partT + 32 * (wl * b.nseg);
  Pt acc = ld_pt((o == s_first * (uint32_t)seg ? H : T) + 32 * (size_t)s_first);
  for (uint32_t s2 = s_first + 1; s2 <= s_last; s2++) acc = pt_add_ni(acc, ld_pt(H + 32 * (size_t)s2));
  return acc;
}

// warp-wide point sum: lane values -> lane 0 (shuffle tree, 5 additions deep)
__device__ __forceinline__ Pt warp_sum_pt(Pt v) {
#pragma unroll 1
  for (int d = 16; d >= 1; d >>= 1) {
    Pt o;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      o.X.w[k] = __shfl_down_sync(0xffffffffu, v.X.w[k], d);
      o.Y.w[k] = __shfl_down_sync(0xffffffffu, v.Y.w[k], d);
      o.Z.w[k] = __shfl_down_sync(0xffffffffu, v.Z.w[k], d);
      o.T.w[k] = __shfl_down_sync(0xffffffffu, v.T.w[k], d);
    }
    v = pt_add_ni(v, o);
  }
  return v;
}

__global__ void __launch_bounds__(128, 5) msm_heavy_kernel(const uint32_t* __restrict__ offs, const uint32_t* __restrict__ hist,
                                                        int nseg, int seg, int nb, const uint32_t* __restrict__ partH,
                                                        const uint32_t* __restrict__ partT, uint32_t* __restrict__ buckets,
                                                        const uint32_t* __restrict__ heavy_count, const uint32_t* __restrict__ heavy_list) {
  const uint32_t nheavy = *heavy_count;
  const int lane = threadIdx.x & 31;
  for (uint32_t h = blockIdx.x * 4 + (threadIdx.x >> 5); h < nheavy; h += gridDim.x * 4) {
    const uint32_t g = heavy_list[h];
    const size_t wl = g / (uint32_t)nb;
    const uint32_t o = offs[g], e = o + hist[g];
    const uint32_t s_first = o / (uint32_t)seg, s_last = (e - 1) / (uint32_t)seg;
    const uint32_t* H = partH + 32 * (wl * nseg);
    const uint32_t* T = partT + 32 * (wl * nseg);
    Pt acc = pt_identity_mont();
    bool have = false;
    for (uint32_t s = s_first + lane; s <= s_last; s += 32) {
      Pt v = ld_pt(((s == s_first && o != s_first * (uint32_t)seg) ? T : H) + 32 * (size_t)s);
      if (!have) { acc = v; have = true; } else acc = pt_add_ni(acc, v);
    }
    acc = warp_sum_pt(acc);
    if (lane == 0) st_pt(buckets + 32 * (size_t)g, acc);
  }
}

__device__ __forceinline__ Pt shfl_down_pt(const Pt& v, int d) {
  Pt o;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    o.X.w[k] = __shfl_down_sync(0xffffffffu, v.X.w[k], d);
    o.Y.w[k] = __shfl_down_sync(0xffffffffu, v.Y.w[k], d);
    o.Z.w[k] = __shfl_down_sync(0xffffffffu, v.Z.w[k], d);
    o.T.w[k] = __shfl_down_sync(0xffffffffu, v.T.w[k], d);
  }
  return o;
}

// ---- short windows: sum the 2^sub sub-buckets of every digit back into one bucket ---------------------------------
// one warp per real bucket k < nb >> sub:  out[k] = sum_t buckets[(k << sub) | t]
__global__ void __launch_bounds__(128) msm_fold_kernel(BucketSrc src, size_t wl, int sub, uint32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const size_t k = (size_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (k >= (size_t)(src.nb >> sub)) return;
  const size_t base = wl * src.nb + (k << sub);
  Pt acc = pt_identity_mont();
  bool have = false;
  for (int t = lane; t < (1 << sub); t += 32) {
    Pt x = load_bucket(src, base + t);
    if (!have) { acc = x; have = true; } else acc = pt_add_ni(acc, x);
  }
  acc = warp_sum_pt(acc);
  if (lane == 0) st_pt(out + 32 * k, acc);
}
// buckets[k] = k < nreal ? folded[k] : identity
__global__ void __launch_bounds__(256) msm_unfold_kernel(const uint32_t* __restrict__ folded, int nb, int nreal, uint32_t* __restrict__ buckets) {
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k >= nb) return;
  Pt p = k < nreal ? ld_pt(folded + 32 * (size_t)k) : pt_identity_mont();
  st_pt(buckets + 32 * (size_t)k, p);
}

// ---- bucket reduction  W = sum_k (k+1) B_k  by digit marginals ("cube" reduction) ----------------------------------
// Write the bucket index as four digits  k = k3 2^(A0+a1+a2) + k2 2^(A0+a1) + k1 2^A0 + k0  (k0: lane, k1: warp in
// block, k2 / k3: low / high bits of the block index).  Then
//     sum_k (k+1) B_k = 2^(A0+a1+a2) sum_v v M3[v] + 2^(A0+a1) sum_v v M2[v] + 2^A0 sum_v v M1[v] + sum_v (v+1) M0[v]
// where Md[v] is the plain sum of all buckets whose digit d equals v.  Plain sums are trees (depth 5 + a1 in stage 1,
// <= 8 in stage 2a), every weighted sum has at most 32 terms (stage 2b, depth ~13) and the powers of two are applied
// for free by the window chain: ~29 dependent point additions instead of ~110 for a chunked running sum.
constexpr int A0 = 5;

// stage 1: one block per (window, k2): 2^a1 warps x 32 lanes, one bucket per thread.
//   pm1[blk][warp] = sum over lanes,  pm0[blk][lane] = sum over warps,  tot[blk] = sum of the block's buckets
// raw_mask: bit wl set = the buckets[] of local window wl (within the group) are final (written by msm_unfold_kernel).
// Blocks are kept to 128 threads / <= 168 registers so that they fit next to the accumulation's resident blocks.
__global__ void __launch_bounds__(128, 3) msm_cube1_kernel(BucketSrc src, uint32_t raw_mask, int a1, uint32_t* __restrict__ tot,
                                                        uint32_t* __restrict__ pm1, uint32_t* __restrict__ pm0) {
  extern __shared__ uint4 cube_sm[];           // [8 uint4 of a point][warp][lane]  +  [8][warp] row sums
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = 1 << a1;
  const size_t blk = blockIdx.x;
  uint4* colbuf = cube_sm;
  uint4* rowbuf = cube_sm + 8 * nw * 32;
  const size_t gidx = blk * (size_t)(nw * 32) + threadIdx.x;
  Pt v = ((raw_mask >> (int)(gidx / src.nb)) & 1u) ? ld_pt(src.buckets + 32 * gidx) : load_bucket(src, gidx);
  auto put = [&](uint4* base, int stride, int idx, const Pt& p) {
    base[0 * stride + idx] = make_uint4(p.X.w[0], p.X.w[1], p.X.w[2], p.X.w[3]); base[1 * stride + idx] = make_uint4(p.X.w[4], p.X.w[5], p.X.w[6], p.X.w[7]);
    base[2 * stride + idx] = make_uint4(p.Y.w[0], p.Y.w[1], p.Y.w[2], p.Y.w[3]); base[3 * stride + idx] = make_uint4(p.Y.w[4], p.Y.w[5], p.Y.w[6], p.Y.w[7]);
    base[4 * stride + idx] = make_uint4(p.Z.w[0], p.Z.w[1], p.Z.w[2], p.Z.w[3]); base[5 * stride + idx] = make_uint4(p.Z.w[4], p.Z.w[5], p.Z.w[6], p.Z.w[7]);
    base[6 * stride + idx] = make_uint4(p.T.w[0], p.T.w[1], p.T.w[2], p.T.w[3]); base[7 * stride + idx] = make_uint4(p.T.w[4], p.T.w[5], p.T.w[6], p.T.w[7]);
  };
  auto get = [&](const uint4* base, int stride, int idx) {
    Pt p; uint4 q;
    q = base[0 * stride + idx]; p.X.w[0] = q.x; p.X.w[1] = q.y; p.X.w[2] = q.z; p.X.w[3] = q.w;
    q = base[1 * stride + idx]; p.X.w[4] = q.x; p.X.w[5] = q.y; p.X.w[6] = q.z; p.X.w[7] = q.w;
    q = base[2 * stride + idx]; p.Y.w[0] = q.x; p.Y.w[1] = q.y; p.Y.w[2] = q.z; p.Y.w[3] = q.w;
    q = base[3 * stride + idx]; p.Y.w[4] = q.x; p.Y.w[5] = q.y; p.Y.w[6] = q.z; p.Y.w[7] = q.w;
    q = base[4 * stride + idx]; p.Z.w[0] = q.x; p.Z.w[1] = q.y; p.Z.w[2] = q.z; p.Z.w[3] = q.w;
    q = base[5 * stride + idx]; p.Z.w[4] = q.x; p.Z.w[5] = q.y; p.Z.w[6] = q.z; p.Z.w[7] = q.w;
    q = base[6 * stride + idx]; p.T.w[0] = q.x; p.T.w[1] = q.y; p.T.w[2] = q.z; p.T.w[3] = q.w;
    q = base[7 * stride + idx]; p.T.w[4] = q.x; p.T.w[5] = q.y; p.T.w[6] = q.z; p.T.w[7] = q.w;
    return p;
  };
  put(colbuf, nw * 32, warp * 32 + lane, v);
  Pt r = warp_sum_pt(v);                       // row sum (lane 0)
  if (lane == 0) { st_pt(pm1 + 32 * (blk * nw + warp), r); put(rowbuf, nw, warp, r); }
  __syncthreads();
  // column sums over the warps (warp 0 ends up with them) and, on the last warp, the block total from the row sums
  Pt t = pt_identity_mont();
  if (warp == nw - 1 && lane < nw) t = get(rowbuf, nw, lane);
  for (int s2 = nw >> 1; s2 >= 1; s2 >>= 1) {
    if (warp < s2) {
      v = pt_add_ni(v, get(colbuf, nw * 32, (warp + s2) * 32 + lane));
      if (s2 > 1) put(colbuf, nw * 32, warp * 32 + lane, v);
    } else if (warp == nw - 1) {
      Pt o = shfl_down_pt(t, s2);
      t = pt_add_ni(t, o);
    }
    __syncthreads();
  }
  if (nw == 1) t = r;                          // single warp: total = its row sum
  if (warp == 0) st_pt(pm0 + 32 * (blk * 32 + lane), v);
  if (warp == nw - 1 && lane == 0) st_pt(tot + 32 * blk, t);
}

// Quad helpers for the reduction kernels: a point lives in four consecutive lanes (lane q holds coordinate q), the
// second operand of an addition is read in full by all four lanes (shared or global memory, a broadcast).  One addition
// costs three multiplication latencies instead of nine -- these kernels are a few warps of dependent additions each.
__device__ __forceinline__ Fe ld_coord(const uint32_t* __restrict__ p, int q) { Fe a; ld_fe(p + 8 * q, a); return a; }
__device__ __forceinline__ void st_coord(uint32_t* __restrict__ p, int q, const Fe& a) { st_fe(p + 8 * q, a); }
__device__ __forceinline__ Fe identity_coord(int q) {          // (0, 1, 1, 0) in Montgomery form
  Fe z{{0, 0, 0, 0, 0, 0, 0, 0}};
  return (q == 1 || q == 2) ? Consts<ModP>::R1() : z;
}
// Tree sum over the quads j < n of a block (n a power of two <= blockDim / 4): the result is in quad 0.  sm: n points.
// Whole warps drop out as the tree narrows (the guard is warp-uniform, quad_add shuffles with a full mask).
__device__ __forceinline__ Fe quad_block_tree(Fe acc, int n, int j, int q, int qbase, uint32_t* __restrict__ sm) {
  const int warp = threadIdx.x >> 5;
  for (int s = n >> 1; s >= 1; s >>= 1) {
    if (j < 2 * s) st_coord(sm + 32 * j, q, acc);
    __syncthreads();
    if (8 * warp < s) {
      const bool on = j < s;
      Fe r = quad_add(acc, ld_pt(sm + 32 * (on ? j + s : j)), q, qbase);
      if (on) acc = r;
    }
    __syncthreads();
  }
  return acc;
}

// stage 2a: one block of C2A_QUADS quads per marginal sum.  Tasks of a window: M1[2^a1] and M0[32] over the window's
// nblk blocks, then M2[2^a2] and M3[2^a3] over the block totals (block index = k3 2^a2 + k2).  marg: [M1 | M0 | M2 | M3].
// C2A_QUADS = 64 when the reduction has the GPU to itself (last group, fixed-base path); 32 for a group whose reduction runs
// beside the next group's accumulation: a full wave of accumulation CTAs (4 x 128 threads x 104 registers per SM) leaves
// 12 K registers per SM -- a 128-thread block of this kernel fits, a 256-thread block waits for the accumulation to END
// (measured at 8 ranks, rank 0: the 240-doubling chain behind it started 140 us late).
template <int C2A_QUADS>
__global__ void __launch_bounds__(4 * C2A_QUADS) msm_cube2a_kernel(const uint32_t* __restrict__ tot, const uint32_t* __restrict__ pm1,
                                                                   const uint32_t* __restrict__ pm0, int a1, int a2, int a3, int nwl,
                                                                   uint32_t* __restrict__ marg) {
  __shared__ __align__(16) uint32_t sm[C2A_QUADS * 32];
  const int q = threadIdx.x & 3, qbase = threadIdx.x & 28, j = threadIdx.x >> 2;
  const int nw = 1 << a1, n2 = 1 << a2, n3 = 1 << a3, nblk = n2 * n3, ntask = nw + 32 + n2 + n3;
  const size_t g = blockIdx.x;
  const size_t wl = g / ntask;
  const int task = (int)(g - wl * ntask);
  const uint32_t* base; size_t stride; int count;
  if (task < nw)                { base = pm1 + 32 * ((wl * nblk) * nw + task);        stride = nw; count = nblk; }
  else if (task < nw + 32)      { base = pm0 + 32 * ((wl * nblk) * 32 + (task - nw)); stride = 32; count = nblk; }
  else if (task < nw + 32 + n2) { base = tot + 32 * (wl * nblk + (task - nw - 32));   stride = n2; count = n3; }
  else                          { base = tot + 32 * (wl * nblk + (size_t)(task - nw - 32 - n2) * n2); stride = 1; count = n2; }
  Fe acc = j < count ? ld_coord(base + 32 * ((size_t)j * stride), q) : identity_coord(q);
  for (int b0 = C2A_QUADS; b0 < count; b0 += C2A_QUADS) {       // count is a power of two: all quads stay busy
    const int b = b0 + j;
    acc = quad_add(acc, ld_pt(base + 32 * ((size_t)b * stride)), q, qbase);
  }
  const int n = count < C2A_QUADS ? count : C2A_QUADS;
  acc = quad_block_tree(acc, n, j, q, qbase, sm);
  if (j == 0) st_coord(marg + 32 * (wl * ntask + task), q, acc);
}

// stage 2b: one block of 32 quads per component: comp[0] = sum v M3[v], comp[1] = sum v M2[v], comp[2] = sum v M1[v],
// comp[3] = sum (v+1) M0[v].   sum_v v I_v = sum_{l >= 1} S_l  with the suffix sums S_l = sum_{v >= l} I_v: a scan and a
// tree, 2 log2(n) additions deep.
// drop_wl: local window (within the launch) whose lane digit k0 is a sub-bucket index (weight (k >> A0) + 1), or -1
__global__ void __launch_bounds__(128) msm_cube2b_kernel(const uint32_t* __restrict__ marg, int a1, int a2, int a3, int drop_wl, uint32_t* __restrict__ comp) {
  __shared__ __align__(16) uint32_t sm[32 * 32];
  const int q = threadIdx.x & 3, qbase = threadIdx.x & 28, j = threadIdx.x >> 2, warp = threadIdx.x >> 5;
  const int nw = 1 << a1, n2 = 1 << a2, n3 = 1 << a3, ntask = nw + 32 + n2 + n3;
  const size_t wl = blockIdx.x >> 2;
  const int which = blockIdx.x & 3;
  const uint32_t* m = marg + 32 * (wl * ntask);
  int n;
  if (which == 0)      { m += 32 * (nw + 32 + n2); n = n3; }
  else if (which == 1) { m += 32 * (nw + 32);      n = n2; }
  else if (which == 2) {                           n = nw; }
  else                 { m += 32 * nw;             n = 32; }
  Fe S = j < n ? ld_coord(m + 32 * j, q) : identity_coord(q);
  for (int d = 1; d < n; d <<= 1) {                              // suffix scan over the quads j < n
    if (j < n) st_coord(sm + 32 * j, q, S);
    __syncthreads();
    if (8 * warp < n) {
      const bool on = j + d < n;
      Fe r = quad_add(S, ld_pt(sm + 32 * (on ? j + d : j)), q, qbase);
      if (on) S = r;
    }
    __syncthreads();
  }
  const Fe total = S;                                            // quad 0: sum of all items
  Fe acc = (j >= 1 && j < n) ? S : identity_coord(q);
  acc = quad_block_tree(acc, n, j, q, qbase, sm);
  if (which == 3) {
    if (j == 0) st_coord(sm, q, total);
    __syncthreads();
    if (warp == 0) {
      Fe r = quad_add(acc, ld_pt(sm), q, qbase);
      acc = ((int)wl == drop_wl) ? total : r;
    }
  }
  if (j == 0) st_coord(comp + 32 * (wl * 4 + which), q, acc);
}

// Fixed-base path, stage 0: make buckets[] final -- one thread per bucket stitches the partials of a bucket that spans a
// few segments and writes the identity into empty ones.  (This is the bulk of the reduction's work, throughput matters:
// one lane per bucket, 38 us for 2^15 buckets of ~5 partials; four lanes per bucket took 52 us.)
__global__ void __launch_bounds__(128) msm_stitch_kernel(BucketSrc src, uint32_t* __restrict__ buckets, size_t total) {
  const size_t g = (size_t)blockIdx.x * 128 + threadIdx.x;
  if (g >= total) return;
  const uint32_t cnt = src.hist[g];
  if (cnt == 0) { if ((long long)src.offs[g] > src.lim_lo && (long long)src.offs[g] <= src.lim_hi) st_pt(buckets + 32 * g, pt_identity_mont()); return; }
  const uint32_t o = src.offs[g], e = o + cnt;
  if ((long long)e <= src.lim_lo || (long long)e > src.lim_hi) return;                   // another launch's bucket (split accumulation)
  const uint32_t s_first = o / (uint32_t)src.seg, s_last = (e - 1) / (uint32_t)src.seg;
  if (s_first == s_last || s_last - s_first + 1 > (uint32_t)src.fix_inline) return;     // already final
  st_pt(buckets + 32 * g, load_bucket(src, g));
}
// The same pass with four lanes per bucket (one quad stitches one bucket, 32 buckets per block): a stitch is a handful of
// dependent additions, two multiplication latencies each on a quad instead of nine in one thread, and a lane holds one
// coordinate + the operand, so the blocks are light enough to run beside the next window's accumulation (whose CTAs
// leave ~24 K registers free on half of the SMs; the one-lane kernel's 256 blocks went through those slots in three
// waves: 97 us against 25 us alone).  The loop runs to the longest stitch of the warp (the shuffles need all lanes).
__global__ void __launch_bounds__(128) msm_stitch_quad_kernel(BucketSrc src, uint32_t* __restrict__ buckets, size_t total) {
  const int q = threadIdx.x & 3, qbase = threadIdx.x & 28;
  const size_t g = (size_t)blockIdx.x * 32 + (threadIdx.x >> 2);
  bool valid = g < total;
  const uint32_t cnt = valid ? src.hist[g] : 0u;
  const uint32_t o = valid ? src.offs[g] : 0u, e = o + cnt;
  valid = valid && (long long)e > src.lim_lo && (long long)e <= src.lim_hi;
  const uint32_t s_first = o / (uint32_t)src.seg, s_last = cnt ? (e - 1) / (uint32_t)src.seg : s_first;
  const size_t wl = valid ? g / src.nb : 0;
  const uint32_t* H = src.partH + 32 * (wl * src.nseg);
  const uint32_t* T = src.partT + 32 * (wl * src.nseg);
  int np = 0;                                                       // partials to add to the first one
  bool write = valid && cnt == 0;                                   // empty bucket: the identity
  Fe acc = identity_coord(q);
  if (valid && cnt != 0 && s_first != s_last && s_last - s_first + 1 <= (uint32_t)src.fix_inline) {
    np = (int)(s_last - s_first);
    write = true;
    acc = ld_coord((o == s_first * (uint32_t)src.seg ? H : T) + 32 * (size_t)s_first, q);
  }
  int mx = np;
#pragma unroll
  for (int d = 16; d >= 4; d >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
#pragma unroll 1
  for (int k = 1; k <= mx; k++) {
    const bool on = k <= np;
    Pt nxt = pt_identity_mont();                                    // quads that are done add the identity (and keep acc): no load of a slot nobody wrote
    if (on) nxt = ld_pt(H + 32 * (size_t)(s_first + k));
    const Fe r = quad_add(acc, nxt, q, qbase);
    if (on) acc = r;
  }
  if (write) st_coord(buckets + 32 * g, q, fe_canon4(acc));          // canonical: load_bucket / msm_fold_kernel read these too
}
// stage 1 with four lanes per point (fixed-base path: one bucket set, nothing else on the GPU at that point; the trees
// are depth-bound).  One block of 64 quads per 128 final buckets (k1 = row 0..3, k0 = column 0..31).  Every quad starts
// in a row tree (16 quads per row); the quads that drop out of it first take the 32 column sums while quad 0 adds up
// the block total.
__global__ void __launch_bounds__(256) msm_cube1_quad_kernel(const uint32_t* __restrict__ buckets, uint32_t* __restrict__ tot,
                                                             uint32_t* __restrict__ pm1, uint32_t* __restrict__ pm0,
                                                             const uint32_t* __restrict__ offs, const uint32_t* __restrict__ hist,
                                                             long long lim_lo, long long lim_hi) {
  __shared__ __align__(16) uint32_t sv[128 * 32];               // the block's buckets
  __shared__ __align__(16) uint32_t sw[64 * 32];                // row trees
  const int q = threadIdx.x & 3, qbase = threadIdx.x & 28, j = threadIdx.x >> 2, warp = threadIdx.x >> 5;
  const size_t blk = blockIdx.x;
  if (offs) {                                                   // split accumulation: the block goes with the launch that completes its last bucket
    const long long e = (long long)offs[blk * 128 + 127] + hist[blk * 128 + 127];
    if (e <= lim_lo || e > lim_hi) return;
  }
  {
    const uint4* src = reinterpret_cast<const uint4*>(buckets + 32 * (blk * 128));
    uint4* dst = reinterpret_cast<uint4*>(sv);
    for (int k = threadIdx.x; k < 128 * 8; k += 256) dst[k] = src[k];
  }
  __syncthreads();
  const int row = j >> 4, i = j & 15;
  Fe acc = quad_add(ld_coord(sv + 32 * (row * 32 + i), q), ld_pt(sv + 32 * (row * 32 + 16 + i)), q, qbase);
  for (int s = 8; s >= 1; s >>= 1) {
    if (i < 2 * s) st_coord(sw + 32 * j, q, acc);
    __syncthreads();
    if ((i & 8) == 0) {                                          // even warps: i = 0..7 of a row
      const bool on = i < s;
      const Fe t = quad_add(acc, ld_pt(sw + 32 * (on ? j + s : j)), q, qbase);
      if (on) acc = t;
    }
    __syncthreads();
  }
  if (i == 0) {                                                  // row sums
    st_coord(pm1 + 32 * (blk * 4 + row), q, acc);
    st_coord(sw + 32 * row, q, acc);
  }
  __syncthreads();
  if (warp == 0) {                                               // quad 0 (row 0): block total
    for (int k = 1; k < 4; k++) acc = quad_add(acc, ld_pt(sw + 32 * k), q, qbase);
    if (j == 0) st_coord(tot + 32 * blk, q, acc);
  } else if (warp & 1) {                                         // the 32 quads with i >= 8: one column each
    const int col = (warp >> 1) * 8 + (i & 7);
    Fe cs = ld_coord(sv + 32 * col, q);
    for (int k = 1; k < 4; k++) cs = quad_add(cs, ld_pt(sv + 32 * (k * 32 + col)), q, qbase);
    st_coord(pm0 + 32 * (blk * 32 + col), q, cs);
  }
}

// The same stage in 128-thread blocks (32 quads) for a reduction that runs BESIDE the next group's accumulation: a full wave
// of accumulation CTAs leaves 12 K registers per SM, so a 256-thread block of msm_cube1_quad_kernel waits until the
// accumulation ends (rank 5 of 8: its high window's chain started 110 us late).  Warp w takes row w: every quad sums four
// buckets of the row, the warp's eight quads finish the row in a three-level tree; then quad j sums column j, and quad 0
// adds up the four row sums.  12 dependent additions instead of 8 -- used only where it overlaps.
__global__ void __launch_bounds__(128) msm_cube1_quad32_kernel(const uint32_t* __restrict__ buckets, uint32_t* __restrict__ tot,
                                                               uint32_t* __restrict__ pm1, uint32_t* __restrict__ pm0) {
  __shared__ __align__(16) uint32_t sv[128 * 32];               // the block's buckets
  __shared__ __align__(16) uint32_t sw[32 * 32];                // row trees, then the four row sums
  const int q = threadIdx.x & 3, qbase = threadIdx.x & 28, j = threadIdx.x >> 2, row = threadIdx.x >> 5, qi = j & 7;
  const size_t blk = blockIdx.x;
  {
    const uint4* src = reinterpret_cast<const uint4*>(buckets + 32 * (blk * 128));
    uint4* dst = reinterpret_cast<uint4*>(sv);
    for (int k = threadIdx.x; k < 128 * 8; k += 128) dst[k] = src[k];
  }
  __syncthreads();
  Fe acc = ld_coord(sv + 32 * (row * 32 + 4 * qi), q);
  for (int t = 1; t < 4; t++) acc = quad_add(acc, ld_pt(sv + 32 * (row * 32 + 4 * qi + t)), q, qbase);
  for (int s = 4; s >= 1; s >>= 1) {
    if (qi < 2 * s) st_coord(sw + 32 * j, q, acc);
    __syncwarp();
    const bool on = qi < s;
    const Fe t = quad_add(acc, ld_pt(sw + 32 * (on ? j + s : j)), q, qbase);
    if (on) acc = t;
    __syncwarp();
  }
  if (qi == 0) st_coord(pm1 + 32 * (blk * 4 + row), q, acc);
  __syncthreads();                                               // every warp is done with its part of sw
  if (qi == 0) st_coord(sw + 32 * row, q, acc);
  Fe cs = ld_coord(sv + 32 * j, q);                              // column j
  for (int k = 1; k < 4; k++) cs = quad_add(cs, ld_pt(sv + 32 * (k * 32 + j)), q, qbase);
  st_coord(pm0 + 32 * (blk * 32 + j), q, cs);
  __syncthreads();
  if (row == 0) {                                                // quad 0: block total from the four row sums
    Fe t = ld_coord(sw, q);
    for (int k = 1; k < 4; k++) t = quad_add(t, ld_pt(sw + 32 * k), q, qbase);
    if (j == 0) st_coord(tot + 32 * blk, q, t);
  }
}

// ---- window chain (four lanes per point operation, see the quad helpers above) ---------------------------------------
// One warp.  acc (in/out, extended Montgomery words) is the running sum, already scaled to 2^(A0+a1+a2) times the unit
// of this group's top window.  Each window contributes four components (msm_cube2b):
//   comp0 2^(A0+a1+a2) + comp1 2^(A0+a1) + comp2 2^A0 + comp3.
//   for i in 0..ng-1 (group windows in descending order):
//     if (i > 0) acc = 2^(gaps.pre[i]) acc        (= c * (window distance) - A0 - a1 - a2)
//     acc += comp0;  acc = 2^a2 acc;  acc += comp1;  acc = 2^a1 acc;  acc += comp2;  acc = 2^A0 acc;  acc += comp3
//   acc = 2^gap_post acc          (gap_post already excludes A0 + a1 + a2 when another group follows)
// first != 0: acc starts as the identity.  out52 != nullptr: also store the result in the ABI layout.
// drop0: the first window's comp3 carries weight 1 on every bucket (sub-bucketed short window): A0 - drop0 doublings.
struct ChainGaps { int pre[MAX_WINDOWS]; };     // doublings before window i of the group (i >= 1), beyond the A0 + a1 + a2 inside it
__global__ void __launch_bounds__(32) msm_chain_kernel(const uint32_t* __restrict__ comp, int ng, int first, int a1, int a2, int drop0,
                                                       const ChainGaps gaps, int gap_post, uint32_t* __restrict__ acc_io,
                                                       uint64_t* __restrict__ out52, const uint64_t* __restrict__ corr52) {
  const int lane = threadIdx.x;
  const int q = lane & 3, qbase = lane & ~3;
  Pt a0 = first ? pt_identity_mont() : ld_pt(acc_io);
  Fe c = pt_coord(a0, q);
#pragma unroll 1
  for (int i = 0; i < ng; i++) {
    const uint32_t* cw = comp - 128 * (ptrdiff_t)i;          // windows in descending order
#pragma unroll 1
    for (int part = 0; part < 4; part++) {
      int nd = part == 0 ? (i > 0 ? gaps.pre[i] : 0) : (part == 1 ? a2 : (part == 2 ? a1 : (i == 0 ? A0 - drop0 : A0)));
#pragma unroll 1
      for (int j = 0; j < nd; j++) c = quad_double(c, q, qbase);
      c = quad_add(c, ld_pt(cw + 32 * part), q, qbase);
    }
  }
#pragma unroll 1
  for (int j = 0; j < gap_post; j++) c = quad_double(c, q, qbase);
  // fixed-base spread correction (ABI layout): normal-form words are a Montgomery-form representative of the same point
  if (corr52) c = quad_add(c, pt_load52(corr52), q, qbase);
  c = fe_canon4(c);                                // the quad operations leave coordinates lazily reduced (< 2.5 m)
  Pt r;
  r.X = shfl_fe(c, 0); r.Y = shfl_fe(c, 1); r.Z = shfl_fe(c, 2); r.T = shfl_fe(c, 3);
  if (lane == 0) {
    st_pt(acc_io, r);
    // Montgomery-form words read as normal form are the same point scaled by R: store them directly.
    if (out52) pt_store52(out52, r);
  }
}

// development aid (ZC_MSM_TRACE=2): device-side timestamps between the kernels of the captured graph
__global__ void msm_stamp_kernel(unsigned long long* slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *slot = t;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int32_t zc_msm_run(zc_ctx *ctx, const zc_msm_generators *gens, const uint64_t *points, const uint64_t *scalars, size_t n, int32_t c,
                   int32_t rank, int32_t nranks, bool exchange, uint64_t *out_point_dev) {
  if (ctx->validate && n) {                                       // scalars >= L (or limbs >= 2^52) would land in wrong buckets silently
    int32_t rc;
    if (points && !gens && (rc = zc_validate_dev(ctx, 3, points, n, 0))) return rc;
    if (scalars && (rc = zc_validate_dev(ctx, 2, scalars, gens ? gens->n : n, 0))) return rc;
    if ((rc = zc_validate_finish(ctx, 1))) return rc;
  }
  if (gens) {
    if (gens->owner != ctx) return zc_fail(ctx, ZC_ERR_STATE, "generator handle belongs to another context");
    if (gens->kind == ZC_GEN_FIXED_BASE && (gens->c != c || gens->rank != rank || gens->nranks != nranks))
      return zc_fail(ctx, ZC_ERR_MODE, "fixed-base tables were built for another (window_bits, rank, nranks)");
    n = gens->n;
  }
  if (c < 8 || c > 16) return zc_fail(ctx, ZC_ERR_MODE, "window_bits must be in 8..16");
  if (nranks < 1 || rank < 0 || rank >= nranks) return zc_fail(ctx, ZC_ERR_SIZE, "bad rank / nranks");
  if (n > ((size_t)1 << 31) - 1) return zc_fail(ctx, ZC_ERR_SIZE, "n exceeds 2^31 - 1");
  const int nwin = (256 + c - 1) / c;
  const int nb = 1 << (c - 1);
  // Tasks of this rank, ascending by window: the windows it owns (window_owner) over all n points.  (A task may also be a
  // point range of a window -- msm_digits_kernel honours [p0, p1).  Splitting the low windows between rank pairs, so that
  // no rank's last window needs more than c (nranks/2 - 1) doublings, was measured at 8 ranks and did not pay: the third
  // bucket reduction per rank costs what the shorter chain saves.)
  struct Task { int w; uint32_t p0, p1; };
  Task tasks[MAX_WINDOWS];
  int nwl = 0;
  for (int w = 0; w < nwin; w++) if (window_owner(w, nranks) == rank) tasks[nwl++] = {w, 0u, (uint32_t)n};
  WinMap wmap;
  for (int w = 0; w < MAX_WINDOWS; w++) { wmap.tl[w] = -1; wmap.p0[w] = 0; wmap.p1[w] = 0; }
  for (int t = 0; t < nwl; t++) { wmap.tl[tasks[t].w] = (int16_t)t; wmap.p0[tasks[t].w] = tasks[t].p0; wmap.p1[tasks[t].w] = tasks[t].p1; }
  wmap.merged = 0;

  uint64_t *partial = exchange ? nullptr : out_point_dev;
  if (exchange) ctx->peer_seq++;                                  // counted before anything can fail: the ranks' sequence numbers never drift apart
  if (exchange) {
    if (!ctx->nccl_comm && !ctx->peers_connected) return zc_fail(ctx, ZC_ERR_STATE, "zc_msm_sharded_dev needs zc_peer_mailbox_connect or zc_ctx_set_nccl first");
    partial = (uint64_t*)ctx->gather_buf + 20 * (size_t)nranks;   // send slot after the nranks receive slots
  }

  if (nwl == 0 || n == 0) {
    // this rank owns no window (nranks > nwin) or the MSM is empty: contribute the identity
    int32_t rc = zc_point_fold_dev(ctx, nullptr, 0, partial);
    if (rc) return rc;
  } else {
    // fixed-base tables prepared for exactly this call shape (zc_msm_prepare_fixed_base_dev): every (window, point) entry is
    // a row of the table, all of the rank's windows share one bucket set, and there is no doubling chain
    const bool use_fb = gens && gens->kind == ZC_GEN_FIXED_BASE;
    const int nwb = use_fb ? 1 : nwl;                           // bucket sets
    wmap.merged = use_fb ? 1 : 0;
    const size_t fb_entries = (size_t)nwl * n;
    // Segment length of an accumulation launch over `entries` sorted entries.  Default by amount of work (32 / 16 / 8); but a
    // launch that can go out as ONE full wave of the 4 x 128-thread slots per SM does: it runs as long as its busiest SM, so
    // 2^20 entries are 586 CTAs of 14-entry segments (four on every SM) rather than 512 CTAs of 16 (68 SMs with four, 80 with
    // three), and 2^21 entries 585 CTAs of 28 rather than 1024 of 16 (measured per rank of eight: 0.557 -> 0.533 ms prepared,
    // 0.392 -> 0.389 ms with tables).  ZC_MSM_SEG forces a length (even, 8..32).
    static const int seg_env = getenv("ZC_MSM_SEG") ? atoi(getenv("ZC_MSM_SEG")) : 0;
    auto pick_seg = [&](size_t entries) -> int {
      if (seg_env >= 8 && seg_env <= SEG_MAX && (seg_env & 1) == 0) return seg_env;
      int seg = entries >= ((size_t)1 << 22) ? 32 : (entries > ((size_t)1 << 19) ? 16 : 8);
      const size_t slots = (size_t)4 * ctx->sm_count * 128;
      const int fill = 2 * (int)((entries + 2 * slots - 1) / (2 * slots));          // smallest even length whose grid fits one wave
      if (fill >= 8 && fill <= SEG_MAX) seg = fill;
      return seg;
    };
    const int fb_seg = pick_seg(fb_entries);
    // ZC_MSM_SORT=atomic: the round-1 sort (one global histogram atomic per entry); default: the two-level counting sort
    static const bool sort_atomic = getenv("ZC_MSM_SORT") && !strcmp(getenv("ZC_MSM_SORT"), "atomic");
    // workspace layout
    size_t o = 0;
    size_t o_cached = o; o = align_up(o + (gens ? 0 : n * 128), 256);
    // Segment length, per task group: 32 when the group holds >= 2^22 entries, 16 below that, 8 for <= 2^19 -- less work
    // gets shorter segments so that the accumulation still fills the GPU (one window of 2^20 points in 32-entry segments
    // is 256 CTAs of 32 serial additions each: latency-bound at 190 us).  Shorter segments mean more partials to stitch
    // per bucket (a 64-entry bucket spans 8-9 segments of 8 and would go to the one-warp-per-bucket heavy path), hence
    // not always 8.  The partial-slot arrays are laid out for the shortest segment.
    const size_t n_pad = align_up(use_fb ? fb_entries : n, SEG_MAX);
    const int nseg_alloc = (int)(n_pad / (use_fb ? fb_seg : 8)) + 1;
    size_t o_digits = o;                                        // sort A: digits (4 B per entry); sort B: coarse-sorted entries (8 B)
    o = align_up(o + ((size_t)nwl * n > (size_t)nwb * n_pad ? (size_t)nwl * n : (size_t)nwb * n_pad) * 8 + 256, 256);
    size_t o_sorted = o; o = align_up(o + (size_t)nwb * n_pad * 4 + 256, 256);
    size_t o_partH = o; o = align_up(o + (size_t)nwb * nseg_alloc * 128, 256);
    size_t o_partT = o; o = align_up(o + (size_t)nwb * nseg_alloc * 128, 256);
    size_t o_heavy = o; o = align_up(o + 256 * MAX_GROUPS + (size_t)nwb * nb * 4 * (use_fb ? MAX_GROUPS : 1), 256);
    size_t o_hist = o;   o = align_up(o + (size_t)nwb * nb * 4, 256);
    size_t o_offs = o;   o = align_up(o + (size_t)nwb * nb * 4, 256);
    size_t o_ranks = o;  o = align_up(o + (sort_atomic ? (size_t)nwl * n * 4 : 0), 256);      // sort A only: rank of each entry inside its bucket
    // A spread short window (fixed-base path) only reaches the lower half of the merged bucket range: those bins take 1.5x the
    // average (2^20 points, rank 0 of 8: 12288 entries per 128-bucket bin, over the 9216 a msm_sort_fine block stages in shared
    // memory -> its slow path, 21 instead of 11 us).  Plan the bins for the heavier half.
    // The same holds for a sub-bucketed short window of the per-window paths when the scalars do not reach its top digit bit
    // (250-bit windows of scalars below L ~ 2^249: half of its slots stay empty, the other half hold twice the average).
    bool skewed = false;
    for (int t = 0; t < nwl; t++) skewed |= use_fb ? merged_spread_bits(c, tasks[t].w) > 0 : short_window_sub_bits(c, tasks[t].w) > 0;
    const size_t sort_entries = use_fb ? fb_entries : n;
    const int sort_cb = sort_coarse_bits(c - 1, skewed ? 2 * sort_entries : sort_entries, nwb), sort_fb = c - 1 - sort_cb;
    const int sort_nblk = (int)((n + SORT_TPB - 1) / SORT_TPB) < 4 * ctx->sm_count ? (int)((n + SORT_TPB - 1) / SORT_TPB) : 4 * ctx->sm_count;
    if (sort_nblk > 32 * BSCAN_PER) return zc_fail(ctx, ZC_ERR_STATE, "device has more SMs than the MSM sort plans for");
    size_t o_scount = o; o = align_up(o + ((size_t)nwb << sort_cb) * sort_nblk * 4, 256);
    size_t o_stot = o;   o = align_up(o + ((size_t)nwb << sort_cb) * 4, 256);
    size_t o_sbin = o;   o = align_up(o + ((size_t)nwb << sort_cb) * 4, 256);
    size_t o_buckets = o; o = align_up(o + (size_t)nwb * nb * 128, 256);
    const int bits = c - 1;                                     // nb = 2^bits, bits in 7..15
    const int a1 = 2;                                           // warps per cube block = 2^a1 (bits >= 7)
    const int ab = bits - A0 - a1;                              // block-index bits, split into k2 (a2) and k3 (a3)
    const int a2 = ab < 4 ? ab : 4, a3 = ab - a2;
    const int nblk = 1 << ab;                                   // cube blocks per window
    const int nw1 = 1 << a1;
    const int ntask = nw1 + 32 + (1 << a2) + (1 << a3);
    size_t o_tot = o;  o = align_up(o + (size_t)nwb * nblk * 128, 256);
    size_t o_pm1 = o;  o = align_up(o + (size_t)nwb * nblk * nw1 * 128, 256);
    size_t o_pm0 = o;  o = align_up(o + (size_t)nwb * nblk * 32 * 128, 256);
    size_t o_marg = o; o = align_up(o + (size_t)nwb * ntask * 128, 256);
    size_t o_fold = o; o = align_up(o + (size_t)4 * nb * 128, 256);      // one per side stream
    size_t o_acc = o;  o = align_up(o + 128, 256);
    size_t o_comp = o; o = align_up(o + (size_t)nwb * 4 * 128, 256);
    if (o > ctx->msm_ws_bytes) {
      if (ctx->msm_ws) ZC_CUDA(ctx, cudaFree(ctx->msm_ws));
      ctx->msm_ws = nullptr; ctx->msm_ws_bytes = 0;
      ZC_CUDA(ctx, cudaMalloc(&ctx->msm_ws, o));
      ctx->msm_ws_bytes = o;
    }
    char *ws = (char*)ctx->msm_ws;
    uint32_t *cached = gens ? (uint32_t*)gens->cached : (uint32_t*)(ws + o_cached);
    int32_t *digits = (int32_t*)(ws + o_digits);
    uint32_t *sorted = (uint32_t*)(ws + o_sorted);
    uint32_t *hist = (uint32_t*)(ws + o_hist);
    uint32_t *partH = (uint32_t*)(ws + o_partH);
    uint32_t *partT = (uint32_t*)(ws + o_partT);
    uint32_t *heavy_count = (uint32_t*)(ws + o_heavy);          // one counter per group, 256 B apart
    uint32_t *heavy_list = heavy_count + 64 * MAX_GROUPS;
    uint32_t *offs = (uint32_t*)(ws + o_offs);
    uint32_t *ranks = (uint32_t*)(ws + o_ranks);
    uint32_t *scount = (uint32_t*)(ws + o_scount), *stot = (uint32_t*)(ws + o_stot), *sbin = (uint32_t*)(ws + o_sbin);
    uint32_t *buckets = (uint32_t*)(ws + o_buckets);
    uint32_t *btot = (uint32_t*)(ws + o_tot);
    uint32_t *pm1 = (uint32_t*)(ws + o_pm1);
    uint32_t *pm0 = (uint32_t*)(ws + o_pm0);
    uint32_t *marg = (uint32_t*)(ws + o_marg);
    uint32_t *folded = (uint32_t*)(ws + o_fold);
    uint32_t *acc = (uint32_t*)(ws + o_acc);
    uint32_t *comp = (uint32_t*)(ws + o_comp);
    cudaStream_t st = ctx->stream;
    if (!ctx->side_stream) {
      ZC_CUDA(ctx, cudaFuncSetAttribute(msm_sort_fine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (FINE_CAP + 2) * 8 + FINE_CAP * 4));
      // highest priority: the side kernels are small and latency-bound; their blocks must not queue behind the
      // accumulation's grid (observed: 4x longer when they do, and the last group's tail waits for them)
      int prio_lo = 0, prio_hi = 0;
      ZC_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
      // the last group's reduction is on the critical path (high priority); the earlier ones have slack and must not take
      // SMs from the accumulation that follows them (low priority)
      ZC_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->side_stream, cudaStreamNonBlocking, prio_lo));
      for (int i = 0; i < 2; i++) ZC_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->side_extra[i], cudaStreamNonBlocking, prio_lo));
      ZC_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->side_extra[2], cudaStreamNonBlocking, prio_hi));
      for (int i = 0; i < 3; i++) ZC_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->side_hi[i], cudaStreamNonBlocking, prio_hi));
      ZC_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->chain_stream, cudaStreamNonBlocking, prio_hi));
      ZC_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->sort_stream, cudaStreamNonBlocking, prio_lo));
      ZC_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->sort_hi, cudaStreamNonBlocking, prio_hi));
      ZC_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->acc2, cudaStreamNonBlocking, prio_lo));
      for (int i = 0; i < 16; i++) ZC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev[i], cudaEventDisableTiming));
    }
    // st: digits, sort, accumulation.  sides[g % 4]: operand preparation (sides[0]), then stitch + reduce of group g --
    // the groups' reductions are independent and latency-bound, so they get their own streams and overlap.  chain: the
    // serial window chain.  Events: ev[0] fork, ev[1] prep done, ev[2+g] group g accumulated, ev[6+g] group g reduced,
    // ev[10] chain done.
    // One GPU: the early groups' reductions have slack (low priority: they must not take SMs from the accumulation that
    // follows).  Sharded: a rank owns few windows and every reduction feeds the serial window chain -- all high priority.
    // accumulation block size (see ACC_TPB_MAX): 64 threads when 128-thread blocks would not fill two waves of the 4-per-SM slots
    static const int acc_tpb_env = getenv("ZC_MSM_ACC_TPB") ? atoi(getenv("ZC_MSM_ACC_TPB")) : 0;
    auto acc_tpb = [&](size_t nthreads) -> int {
      if (acc_tpb_env == 64 || acc_tpb_env == 128) return acc_tpb_env;
      (void)nthreads;
      return 128;                                 // measured at 8 ranks (one window of 2^20 entries per launch): 64-thread blocks change nothing (176 us either way)
    };
    static const bool stitch_quad = !(getenv("ZC_MSM_STITCH_QUAD") && atoi(getenv("ZC_MSM_STITCH_QUAD")) == 0);
    static const int hi_prio_env = getenv("ZC_MSM_HI_PRIO") ? atoi(getenv("ZC_MSM_HI_PRIO")) : -1;
    const bool hi_prio = hi_prio_env >= 0 ? hi_prio_env != 0 : nranks > 1;
    cudaStream_t sides[4] = {hi_prio ? ctx->side_hi[0] : ctx->side_stream, hi_prio ? ctx->side_hi[1] : ctx->side_extra[0],
                             hi_prio ? ctx->side_hi[2] : ctx->side_extra[1], ctx->side_extra[2]};
    cudaStream_t side = sides[0], chain = ctx->chain_stream;

    // The ~25 launches and the two-stream fork/join of one MSM are recorded once into a CUDA graph and replayed while
    // the call's arguments stay the same (repeated proofs over resident generators): one launch instead of a launch-
    // latency-bound sequence -- at 8 ranks the per-rank kernels are short enough for launch gaps to rival the math.
    // generator handle (zc_msm_generators_create_dev): the operand pass ran once, the handle owns the cached array
    const bool use_prepared = gens && gens->kind == ZC_GEN_PREPARED;
    uint64_t nlaunch = 0;
    // ZC_MSM_TRACE=1: no graph, a timing event after every kernel, timeline printed to stderr (development aid).
    // ZC_MSM_TRACE=2: the graph as usual, with a one-thread %globaltimer stamp after every kernel; timeline printed
    // after each call (synchronises).
    static const int trace_mode = getenv("ZC_MSM_TRACE") ? atoi(getenv("ZC_MSM_TRACE")) : 0;
    const bool trace = trace_mode == 1;
    struct Mark { cudaEvent_t ev; const char* name; int stream; };
    std::vector<Mark> marks;
    static unsigned long long* stamp_buf = nullptr;
    static std::vector<std::pair<const char*, int>> stamp_names;
    if (trace_mode == 2 && !stamp_buf) cudaMalloc(&stamp_buf, 256 * sizeof(unsigned long long));
    int nstamp = 0;
    auto mark = [&](cudaStream_t s_, int sid, const char* name) {
      if (trace_mode == 2) {
        if (nstamp == 0) stamp_names.clear();
        if (nstamp < 256) { msm_stamp_kernel<<<1, 1, 0, s_>>>(stamp_buf + nstamp); if ((int)stamp_names.size() <= nstamp) stamp_names.push_back({name, sid}); nstamp++; }
        return;
      }
      if (!trace) return;
      cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s_); marks.push_back({e, name, sid});
    };
    auto enqueue = [&]() -> int32_t {
      mark(st, 0, "start");
      if (sort_atomic) ZC_CUDA(ctx, cudaMemsetAsync(hist, 0, (size_t)nwb * nb * 4, st));
      ZC_CUDA(ctx, cudaMemsetAsync(heavy_count, 0, 256 * MAX_GROUPS, st));
      if (!use_prepared && !use_fb) {
        // The operand pass streams 288 bytes per point through HBM, the first group's sort is latency-bound: they share the
        // GPU well, but only if the sort's few hundred blocks are not queued behind the pass's 8192 -- the pass goes to the
        // low-priority side stream, the sort (sharded: sort_hi) above it.  Both must finish before the first accumulation.
        cudaStream_t ps = ctx->side_stream;
        ZC_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
        ZC_CUDA(ctx, cudaStreamWaitEvent(ps, ctx->ev[0], 0));
        const size_t prep_blocks = (n + PREP_TPB - 1) / PREP_TPB, prep_cap = (size_t)(nranks > 1 ? 3 : 6) * ctx->sm_count;
        if (((uintptr_t)points & 15) == 0) msm_prep_kernel<<<(unsigned)(prep_blocks < prep_cap ? prep_blocks : prep_cap), PREP_TPB, 0, ps>>>(points, cached, n);
        else msm_prep_simple_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ps>>>(points, cached, n);
        nlaunch++; mark(ps, 1, "msm_prep_kernel");
        ZC_CUDA(ctx, cudaEventRecord(ctx->ev[1], ps));
      }
      // digits + histogram, scan, scatter of the local windows [lo, hi) on stream s_
      auto sort_windows = [&](cudaStream_t s_, int sid, int lo, int hi) {
        WinMap m = wmap;
        for (int w = 0; w < MAX_WINDOWS; w++) if (m.tl[w] < lo || m.tl[w] >= hi) m.tl[w] = -1;
        if (!sort_atomic) {
          SortPlan pl;
          pl.cb = sort_cb; pl.fb = sort_fb; pl.chunk = (uint32_t)((n + sort_nblk - 1) / sort_nblk); pl.lo = use_fb ? 0 : lo;
          const int nprob = use_fb ? 1 : hi - lo, nbins = nprob << sort_cb;
          const size_t po = use_fb ? 0 : (size_t)lo;             // this sort's first bucket set
          uint64_t *tmp = (uint64_t*)digits + po * n_pad;
          switch (c) {
#define ZC_SORT_CASE(C) case C: msm_sort_count_kernel<C><<<sort_nblk, SORT_TPB, nbins * 4, s_>>>(scalars, n, m, pl, nbins, scount); break;
            ZC_SORT_CASE(8) ZC_SORT_CASE(9) ZC_SORT_CASE(10) ZC_SORT_CASE(11) ZC_SORT_CASE(12)
            ZC_SORT_CASE(13) ZC_SORT_CASE(14) ZC_SORT_CASE(15) ZC_SORT_CASE(16)
#undef ZC_SORT_CASE
          }
          nlaunch++; mark(s_, sid, "msm_sort_count_kernel");
          msm_sort_bscan_kernel<<<(unsigned)((nbins + 3) / 4), 128, 0, s_>>>(scount, sort_nblk, nbins, stot); nlaunch++; mark(s_, sid, "msm_sort_bscan_kernel");
          switch (c) {
#define ZC_SORT_CASE(C) case C: msm_sort_place_kernel<C><<<sort_nblk, SORT_TPB, 2 * nbins * 4, s_>>>(scalars, n, n_pad, m, pl, nbins, scount, stot, sbin, tmp); break;
            ZC_SORT_CASE(8) ZC_SORT_CASE(9) ZC_SORT_CASE(10) ZC_SORT_CASE(11) ZC_SORT_CASE(12)
            ZC_SORT_CASE(13) ZC_SORT_CASE(14) ZC_SORT_CASE(15) ZC_SORT_CASE(16)
#undef ZC_SORT_CASE
          }
          nlaunch++; mark(s_, sid, "msm_sort_place_kernel");
          msm_sort_fine_kernel<<<(unsigned)nbins, FINE_TPB, (FINE_CAP + 2) * 8 + FINE_CAP * 4, s_>>>(tmp, sbin, stot, pl, n_pad, nb, hist + po * nb, offs + po * nb, sorted + po * n_pad);
          nlaunch++; mark(s_, sid, "msm_sort_fine_kernel");
          return;
        }
        const unsigned grid = (unsigned)((n + 255) / 256);
        switch (c) {
#define ZC_DIGITS_CASE(C) case C: msm_digits_kernel<C><<<grid, 256, 0, s_>>>(scalars, n, m, digits, ranks, hist); break;
          ZC_DIGITS_CASE(8) ZC_DIGITS_CASE(9) ZC_DIGITS_CASE(10) ZC_DIGITS_CASE(11) ZC_DIGITS_CASE(12)
          ZC_DIGITS_CASE(13) ZC_DIGITS_CASE(14) ZC_DIGITS_CASE(15) ZC_DIGITS_CASE(16)
#undef ZC_DIGITS_CASE
        }
        nlaunch++; mark(s_, sid, "msm_digits_kernel");
        if (use_fb) {
          msm_scan_kernel<<<1, SCAN_TPB, 0, s_>>>(hist, offs, nb); nlaunch++; mark(s_, sid, "msm_scan_kernel");
          msm_scatter_kernel<<<(unsigned)((n * (size_t)nwl + 255) / 256), 256, 0, s_>>>(digits, ranks, n, n_pad, nwl, nb, 1, offs, sorted);
        } else {
          msm_scan_kernel<<<hi - lo, SCAN_TPB, 0, s_>>>(hist + (size_t)lo * nb, offs + (size_t)lo * nb, nb); nlaunch++; mark(s_, sid, "msm_scan_kernel");
          msm_scatter_kernel<<<(unsigned)((n * (size_t)(hi - lo) + 255) / 256), 256, 0, s_>>>(digits + (size_t)lo * n, ranks + (size_t)lo * n, n, n_pad,
              hi - lo, nb, 0, offs + (size_t)lo * nb, sorted + (size_t)lo * n_pad);
        }
        nlaunch++; mark(s_, sid, "msm_scatter_kernel");
      };
      // Task groups, top-down (local index wl ascends with the window index)
      const int ngroups = nwl < MAX_GROUPS ? nwl : MAX_GROUPS;
      int glo[MAX_GROUPS], ghi[MAX_GROUPS];
      // Group sizes: equal.  (Experiment, ZC_MSM_GROUP_SPLIT=1: two pieces of the pipeline are exposed -- the sort of the FIRST
      // group and the reduction + window chain of the LAST one -- so give the outer groups half the windows of the inner ones,
      // 2, 6, 6, 2 instead of 4, 4, 4, 4 at 16 windows.  Measured at 2^20 points: 2.730 vs 2.666 ms prepared, 3.009 vs 2.957 ms
      // arbitrary points on one GPU, 1.523 vs 1.448 ms for rank 1 of 2 -- the six-window accumulations lose more than the
      // shorter ends gain.  Off.)
      static const bool uneven_env = getenv("ZC_MSM_GROUP_SPLIT") && atoi(getenv("ZC_MSM_GROUP_SPLIT")) == 1;
      int gsizes[MAX_GROUPS];
      for (int g = 0, h = nwl; g < ngroups; g++) { gsizes[g] = (h + (ngroups - g) - 1) / (ngroups - g); h -= gsizes[g]; }
      if (uneven_env && ngroups == 4 && nwl >= 8) {
        const int k = nwl / 8, m = (nwl - 2 * k + 1) / 2;
        gsizes[0] = k; gsizes[1] = m; gsizes[2] = nwl - 2 * k - m; gsizes[3] = k;
      }
      for (int g = 0, h = nwl; g < ngroups; g++) { ghi[g] = h; glo[g] = h - gsizes[g]; h -= gsizes[g]; }
      // The sort of a group is atomics / scattered stores, its accumulation is multiplier-bound: the groups' sorts run on
      // their own stream, one or more groups ahead of the accumulation (2^20 points on one GPU: 0.4 ms of sorting, three
      // quarters of it hidden).
      static const bool pipe_sort_env = !(getenv("ZC_MSM_PIPE_SORT") && atoi(getenv("ZC_MSM_PIPE_SORT")) == 0);
      const bool pipe_sort = pipe_sort_env && !use_fb && ngroups > 1;
      if (pipe_sort) {
        // sharded (few windows per rank, everything on the critical path): high priority; one GPU: low, under the accumulations
        cudaStream_t ss = nranks > 1 ? ctx->sort_hi : ctx->sort_stream;
        ZC_CUDA(ctx, cudaEventRecord(ctx->ev[15], st));
        ZC_CUDA(ctx, cudaStreamWaitEvent(ss, ctx->ev[15], 0));
        for (int g = 0; g < ngroups; g++) {
          sort_windows(ss, 3, glo[g], ghi[g]);
          ZC_CUDA(ctx, cudaEventRecord(ctx->ev[11 + g], ss));
        }
      } else {
        sort_windows(st, 0, 0, nwl);
      }
      if (use_fb) {
        // one bucket set over all (window, point) entries: accumulate, stitch, reduce, combine the four cube components
        // (A0 + a1 + a2 doublings in all).  A typical bucket spans entries / (nb seg) segments; up to twice that is stitched
        // inline by the reduction's loads, the few heavier ones (buckets the short top window also feeds) one warp each.
        const int seg = fb_seg, nseg = (int)((n_pad + seg - 1) / seg);
        const int fix_inline = 2 * (int)(fb_entries / ((size_t)nb * seg)) + FIX_INLINE;
        // Split accumulation (experiment, ZC_MSM_SPLIT=2..4; off by default): the sorted list is accumulated by `parts` launches
        // over consecutive segment ranges (each on its own stream, the block scheduler drains the earlier launch first) and the
        // buckets a range completes are stitched and taken through stage 1 of the reduction on a high-priority side stream
        // while the next range still accumulates.  Measured at 8 ranks, 2^20 points: 0.427 ms unsplit, 0.438 / 0.474 / 0.515 ms
        // with 2 / 3 / 4 parts -- stage 1 is throughput-bound too (it takes from the accumulation what it hides), and the
        // last part's cube1 / cube2a / cube2b / chain are depth-bound: half the buckets do not make them shorter.
        static const int parts_env = getenv("ZC_MSM_SPLIT") ? atoi(getenv("ZC_MSM_SPLIT")) : 0;
        int parts = parts_env >= 1 && parts_env <= MAX_GROUPS ? parts_env : 1;
        if (nseg < parts * 128) parts = 1;
        cudaStream_t acc_st[MAX_GROUPS] = {st, ctx->side_stream, ctx->side_extra[0], ctx->side_extra[1]};
        cudaStream_t tail = parts > 1 ? ctx->side_extra[2] : st;                  // high priority
        const int tid_ = parts > 1 ? 1 : 0;
        if (parts > 1) {
          ZC_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));                          // sorted
          for (int p = 1; p < parts; p++) ZC_CUDA(ctx, cudaStreamWaitEvent(acc_st[p], ctx->ev[0], 0));
        }
        const uint32_t per = (uint32_t)(((nseg + parts - 1) / parts + 127) / 128 * 128);   // segments per part, whole blocks
        for (int p = 0; p < parts; p++) {
          const uint32_t g_lo = (uint32_t)p * per, g_hi = (p == parts - 1 || g_lo + per > (uint32_t)nseg) ? (uint32_t)nseg : g_lo + per;
          if (g_lo < g_hi) {
            if (acc_tpb((size_t)nseg) == 64)
              msm_accum_kernel<true, 64><<<(unsigned)((g_hi - g_lo + 63) / 64), 64, 0, acc_st[p]>>>((const uint32_t*)gens->table, sorted, offs, hist, n_pad, nseg, seg, 1, nb, buckets, partH, partT, g_lo, g_hi);
            else
              msm_accum_kernel<true, 128><<<(unsigned)((g_hi - g_lo + 127) / 128), 128, 0, acc_st[p]>>>((const uint32_t*)gens->table, sorted, offs, hist, n_pad, nseg, seg, 1, nb, buckets, partH, partT, g_lo, g_hi);
            nlaunch++; mark(acc_st[p], p == 0 ? 0 : 3, "msm_accum_kernel");
          }
          if (parts > 1) {
            ZC_CUDA(ctx, cudaEventRecord(ctx->ev[2 + p], acc_st[p]));
            ZC_CUDA(ctx, cudaStreamWaitEvent(tail, ctx->ev[2 + p], 0));
          }
          // entries below g_hi * seg are accumulated: the buckets that end there are complete
          const long long lim_lo = p == 0 ? LIM_ALL_LO : (long long)g_lo * seg;
          const long long lim_hi = p == parts - 1 ? LIM_ALL_HI : (long long)g_hi * seg;
          uint32_t *hc = heavy_count + 64 * p, *hl = heavy_list + (size_t)p * nb;
          msm_fixq_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, tail>>>(offs, hist, seg, 1, nb, fix_inline, hc, hl, lim_lo, lim_hi); nlaunch++; mark(tail, tid_, "msm_fixq_kernel");
          msm_heavy_kernel<<<2 * ctx->sm_count, 128, 0, tail>>>(offs, hist, nseg, seg, nb, partH, partT, buckets, hc, hl); nlaunch++; mark(tail, tid_, "msm_heavy_kernel");
          const BucketSrc src = {offs, hist, partH, partT, buckets, nseg, nb, seg, fix_inline, lim_lo, lim_hi};
          if (stitch_quad) { msm_stitch_quad_kernel<<<(unsigned)((nb + 31) / 32), 128, 0, tail>>>(src, buckets, (size_t)nb); nlaunch++; mark(tail, tid_, "msm_stitch_quad_kernel"); }
          else { msm_stitch_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, tail>>>(src, buckets, (size_t)nb); nlaunch++; mark(tail, tid_, "msm_stitch_kernel"); }
          msm_cube1_quad_kernel<<<(unsigned)nblk, 256, 0, tail>>>(buckets, btot, pm1, pm0, parts > 1 ? offs : nullptr, hist, lim_lo, lim_hi); nlaunch++; mark(tail, tid_, "msm_cube1_quad_kernel");
        }
        msm_cube2a_kernel<64><<<(unsigned)ntask, 256, 0, tail>>>(btot, pm1, pm0, a1, a2, a3, 1, marg); nlaunch++; mark(tail, tid_, "msm_cube2a_kernel");
        msm_cube2b_kernel<<<4, 128, 0, tail>>>(marg, a1, a2, a3, -1, comp); nlaunch++; mark(tail, tid_, "msm_cube2b_kernel");
        ChainGaps gaps;
        for (int i = 0; i < MAX_WINDOWS; i++) gaps.pre[i] = 0;
        msm_chain_kernel<<<1, 32, 0, tail>>>(comp, 1, 1, a1, a2, 0, gaps, 0, acc, partial, (const uint64_t*)gens->corr); nlaunch++; mark(tail, tid_, "msm_chain_kernel");
        if (parts > 1) {
          ZC_CUDA(ctx, cudaEventRecord(ctx->ev[10], tail));
          ZC_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev[10], 0));
        }
        ZC_CUDA(ctx, cudaGetLastError());
        return ZC_OK;
      }
      // Task groups, top-down (local index wl ascends with the window index).  After a group's buckets are
      // accumulated the side stream stitches and reduces them, folds the window sums into acc and scales acc down to the
      // next group's top window (or, after the last group, by 2^(c * rank)) while the main stream accumulates the next group.
      for (int g = 0; g < ngroups; g++) {
        side = (g == ngroups - 1) ? sides[3] : sides[g % 3];
        const int hi = ghi[g], lo = glo[g], gsz = hi - lo;
        // ZC_MSM_ACC_OVERLAP=1 (experiment, off): consecutive groups accumulate on alternating streams, so the next group's CTAs
        // could fill the slots the current group's last, partial wave leaves idle.  Measured: nothing on one GPU (2.659 vs
        // 2.660 ms -- the next group's low-priority sort only finishes when the current accumulation ends, and the machine is
        // throughput-bound: whatever runs beside an accumulation costs it about its own stand-alone time), worse for rank 0 of 8
        // (0.612 vs 0.521 ms: the low window's CTAs delay the high window's reduction and its 240-doubling chain).
        static const bool acc_overlap = getenv("ZC_MSM_ACC_OVERLAP") && atoi(getenv("ZC_MSM_ACC_OVERLAP")) == 1;
        cudaStream_t as = (acc_overlap && pipe_sort && (g & 1)) ? ctx->acc2 : st;
        if (pipe_sort) ZC_CUDA(ctx, cudaStreamWaitEvent(as, ctx->ev[11 + g], 0));                 // this group's entries are sorted
        size_t group_entries = 0;
        for (int wl = lo; wl < hi; wl++) group_entries += tasks[wl].p1 - tasks[wl].p0;
        const int seg = pick_seg(group_entries);
        const int nseg = (int)((n_pad + seg - 1) / seg);
        const size_t tot = (size_t)gsz * nb;
        const size_t tseg = (size_t)gsz * nseg;
        const uint32_t *g_sorted = sorted + (size_t)lo * n_pad, *g_offs = offs + (size_t)lo * nb, *g_hist = hist + (size_t)lo * nb;
        uint32_t *g_buckets = buckets + 32 * ((size_t)lo * nb), *g_partH = partH + 32 * ((size_t)lo * nseg_alloc), *g_partT = partT + 32 * ((size_t)lo * nseg_alloc);
        uint32_t *g_hcount = heavy_count + 64 * g, *g_hlist = heavy_list + (size_t)lo * nb;
        if (g == 0 && !use_prepared) ZC_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev[1], 0));   // cached operands ready
        const int tpb = acc_tpb(tseg);
        if (use_prepared && tpb == 64)
          msm_accum_kernel<true, 64><<<(unsigned)((tseg + 63) / 64), 64, 0, as>>>(cached, g_sorted, g_offs, g_hist, n_pad, nseg, seg, gsz, nb, g_buckets, g_partH, g_partT, 0u, (uint32_t)tseg);
        else if (use_prepared)
          msm_accum_kernel<true, 128><<<(unsigned)((tseg + 127) / 128), 128, 0, as>>>(cached, g_sorted, g_offs, g_hist, n_pad, nseg, seg, gsz, nb, g_buckets, g_partH, g_partT, 0u, (uint32_t)tseg);
        else if (tpb == 64)
          msm_accum_kernel<false, 64><<<(unsigned)((tseg + 63) / 64), 64, 0, as>>>(cached, g_sorted, g_offs, g_hist, n_pad, nseg, seg, gsz, nb, g_buckets, g_partH, g_partT, 0u, (uint32_t)tseg);
        else
          msm_accum_kernel<false, 128><<<(unsigned)((tseg + 127) / 128), 128, 0, as>>>(cached, g_sorted, g_offs, g_hist, n_pad, nseg, seg, gsz, nb, g_buckets, g_partH, g_partT, 0u, (uint32_t)tseg);
        nlaunch++; mark(as, as == st ? 0 : 4, "msm_accum_kernel");
        // everything after the accumulation is latency-bound (few warps, long dependent chains): it runs on the side
        // stream, under the next group's accumulation
        // Stage 1 of a reduction (stitch + the first trees) is wide: beside the next group's accumulation it competes for the
        // multiplier pipe and takes 4-5x longer (97 + 45 us instead of 21 + 25 us for one window of 2^15 buckets).  On one
        // GPU that is hidden; sharded, the first group's reduction feeds a long window chain, so when that chain is long
        // stage 1 runs ON the main stream, ahead of the next accumulation.  ZC_MSM_SEQ_STAGE1 = 0 / 1 forces it.
        static const int seq_env = getenv("ZC_MSM_SEQ_STAGE1") ? atoi(getenv("ZC_MSM_SEQ_STAGE1")) : -1;
        const int chain_after = (g == ngroups - 1) ? 0 : c * (tasks[lo].w - tasks[lo - 1].w);   // doublings this group's sums wait for
        // In-line stage 1 delays the later accumulations by ~45 us and brings this group's sums forward by ~130 us: worth it
        // when the chain that waits for them is longer than ~100 doublings (measured per rank at 8 ranks, 2^20 points).
        const int post_all = c * tasks[0].w;                     // doublings after the last group (the rank's lowest window)
        const bool seq1 = (g < ngroups - 1) && (seq_env >= 0 ? seq_env != 0 : (nranks > 1 && g == 0 && chain_after > 100));
        (void)post_all;
        cudaStream_t s1 = seq1 ? as : side;
        const int s1id = seq1 ? 0 : 1;
        if (!seq1) {
          ZC_CUDA(ctx, cudaEventRecord(ctx->ev[2 + g], as));
          ZC_CUDA(ctx, cudaStreamWaitEvent(side, ctx->ev[2 + g], 0));
        }
        msm_fixq_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s1>>>(g_offs, g_hist, seg, gsz, nb, FIX_INLINE, g_hcount, g_hlist, LIM_ALL_LO, LIM_ALL_HI); nlaunch++; mark(s1, s1id, "msm_fixq_kernel");
        msm_heavy_kernel<<<2 * ctx->sm_count, 128, 0, s1>>>(g_offs, g_hist, nseg, seg, nb, g_partH, g_partT, g_buckets, g_hcount, g_hlist); nlaunch++; mark(s1, s1id, "msm_heavy_kernel");
        const BucketSrc src = {g_offs, g_hist, g_partH, g_partT, g_buckets, nseg, nb, seg, FIX_INLINE, LIM_ALL_LO, LIM_ALL_HI};
        // A short (top) window spreads each digit over 2^sub sub-buckets.  sub == A0: the sub-bucket index is exactly the
        // lane digit of the cube, which then simply carries weight 0 (drop).  Otherwise sum the sub-buckets back first.
        uint32_t raw_mask = 0; int drop_wl = -1;
        // stage 1 either stitches on load (msm_cube1_kernel, one bucket per thread) or, like the fixed-base path, reads buckets
        // made final by a stitch pass and runs its trees with four lanes per point.  ZC_MSM_QUAD_STAGE1: 0 never, 1 every
        // group (default), 2 only the last group.  Measured at 2^20 points on one GPU: 3.40 / 3.23 / 3.30 ms -- the
        // reductions share the SMs with the next group's accumulation, so their total SM time matters, not only their depth.
        static const int quad_stage1_env = getenv("ZC_MSM_QUAD_STAGE1") ? atoi(getenv("ZC_MSM_QUAD_STAGE1")) : 1;
        const bool quad_stage1 = quad_stage1_env == 1 || (quad_stage1_env == 2 && g == ngroups - 1);
        if (quad_stage1 && stitch_quad) { msm_stitch_quad_kernel<<<(unsigned)((tot + 31) / 32), 128, 0, s1>>>(src, g_buckets, tot); nlaunch++; mark(s1, s1id, "msm_stitch_quad_kernel"); }
        else if (quad_stage1) { msm_stitch_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, s1>>>(src, g_buckets, tot); nlaunch++; mark(s1, s1id, "msm_stitch_kernel"); }
        for (int wl = lo; wl < hi; wl++) {
          const int sub = short_window_sub_bits(c, tasks[wl].w);
          if (sub == A0 && g == 0 && wl == hi - 1) drop_wl = wl - lo;
          else if (sub > 0) {
            raw_mask |= 1u << (wl - lo);
            msm_fold_kernel<<<(unsigned)(((nb >> sub) + 3) / 4), 128, 0, s1>>>(src, (size_t)(wl - lo), sub, folded + (size_t)(g == ngroups - 1 ? 3 : g % 3) * nb * 32); nlaunch++; mark(s1, s1id, "msm_fold_kernel");
            msm_unfold_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, s1>>>(folded + (size_t)(g == ngroups - 1 ? 3 : g % 3) * nb * 32, nb, nb >> sub, buckets + 32 * ((size_t)wl * nb)); nlaunch++; mark(s1, s1id, "msm_unfold_kernel");
          }
        }
        const size_t cube_smem = (size_t)(8 * nw1 * 32 + 8 * nw1) * sizeof(uint4);
        if (quad_stage1 && !seq1 && g < ngroups - 1 && nranks > 1) {   // beside the next accumulation (see msm_cube1_quad32_kernel)
          msm_cube1_quad32_kernel<<<(unsigned)((size_t)gsz * nblk), 128, 0, s1>>>(g_buckets, btot + 32 * ((size_t)lo * nblk),
              pm1 + 32 * ((size_t)lo * nblk * nw1), pm0 + 32 * ((size_t)lo * nblk * 32)); nlaunch++; mark(s1, s1id, "msm_cube1_quad32_kernel");
        } else if (quad_stage1) {
          msm_cube1_quad_kernel<<<(unsigned)((size_t)gsz * nblk), 256, 0, s1>>>(g_buckets, btot + 32 * ((size_t)lo * nblk),
              pm1 + 32 * ((size_t)lo * nblk * nw1), pm0 + 32 * ((size_t)lo * nblk * 32), nullptr, nullptr, LIM_ALL_LO, LIM_ALL_HI); nlaunch++; mark(s1, s1id, "msm_cube1_quad_kernel");
        } else {
          msm_cube1_kernel<<<(unsigned)((size_t)gsz * nblk), 32 * nw1, cube_smem, s1>>>(src, raw_mask, a1, btot + 32 * ((size_t)lo * nblk),
              pm1 + 32 * ((size_t)lo * nblk * nw1), pm0 + 32 * ((size_t)lo * nblk * 32)); nlaunch++; mark(s1, s1id, "msm_cube1_kernel");
        }
        if (seq1) {
          ZC_CUDA(ctx, cudaEventRecord(ctx->ev[2 + g], as));
          ZC_CUDA(ctx, cudaStreamWaitEvent(side, ctx->ev[2 + g], 0));
        }
        if (g == ngroups - 1)
          msm_cube2a_kernel<64><<<(unsigned)((size_t)gsz * ntask), 256, 0, side>>>(btot + 32 * ((size_t)lo * nblk),
              pm1 + 32 * ((size_t)lo * nblk * nw1), pm0 + 32 * ((size_t)lo * nblk * 32), a1, a2, a3, gsz, marg + 32 * ((size_t)lo * ntask));
        else
          msm_cube2a_kernel<32><<<(unsigned)((size_t)gsz * ntask), 128, 0, side>>>(btot + 32 * ((size_t)lo * nblk),
              pm1 + 32 * ((size_t)lo * nblk * nw1), pm0 + 32 * ((size_t)lo * nblk * 32), a1, a2, a3, gsz, marg + 32 * ((size_t)lo * ntask));
        nlaunch++; mark(side, 1, "msm_cube2a_kernel");
        msm_cube2b_kernel<<<4 * gsz, 128, 0, side>>>(marg + 32 * ((size_t)lo * ntask), a1, a2, a3, drop_wl, comp + 128 * (size_t)lo); nlaunch++; mark(side, 1, "msm_cube2b_kernel");
        // only the top local window of the whole MSM can be short, i.e. the first window of the first group
        const int drop0 = (drop_wl == gsz - 1) ? A0 : 0;
        ZC_CUDA(ctx, cudaEventRecord(ctx->ev[6 + g], side));
        ZC_CUDA(ctx, cudaStreamWaitEvent(chain, ctx->ev[6 + g], 0));
        const bool last = (g == ngroups - 1);
        ChainGaps gaps;
        for (int i = 0; i < MAX_WINDOWS; i++) gaps.pre[i] = 0;
        for (int i = 1; i < gsz; i++) gaps.pre[i] = c * (tasks[hi - i].w - tasks[hi - 1 - i].w) - A0 - a1 - a2;
        const int gap_post = last ? c * tasks[lo].w : c * (tasks[lo].w - tasks[lo - 1].w) - A0 - a1 - a2;
        msm_chain_kernel<<<1, 32, 0, chain>>>(comp + 128 * (size_t)(hi - 1), gsz, g == 0 ? 1 : 0, a1, a2, drop0, gaps,
                                             gap_post, acc, last ? partial : nullptr, nullptr); nlaunch++; mark(chain, 2, "msm_chain_kernel");
      }
      ZC_CUDA(ctx, cudaEventRecord(ctx->ev[10], chain));
      ZC_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev[10], 0));
      ZC_CUDA(ctx, cudaGetLastError());
      return ZC_OK;
    };
    zc_msm_key key;
    memset(&key, 0, sizeof(key));                               // padding bytes take part in the memcmp below
    key.points = gens ? nullptr : points; key.scalars = scalars; key.n = n; key.c = c; key.rank = rank; key.nranks = nranks;
    key.mode = use_fb ? 2 : (use_prepared ? 1 : 0); key.partial = partial; key.ws = ctx->msm_ws; key.gens_id = gens ? gens->id : 0;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    ZC_CUDA(ctx, cudaStreamIsCapturing(st, &cap));
    if (cap != cudaStreamCaptureStatusNone || trace) {
      int32_t rc = enqueue();                                   // the caller is capturing: become part of their graph
      if (rc) return rc;
      ctx->launches += nlaunch;
      if (trace && !marks.empty()) {
        cudaStreamSynchronize(st);
        fprintf(stderr, "[zc_msm trace] n=%zu c=%d rank %d/%d\n", n, c, rank, nranks);
        for (size_t i = 1; i < marks.size(); i++) {
          float ms = 0; cudaEventElapsedTime(&ms, marks[0].ev, marks[i].ev);
          fprintf(stderr, "  %8.1f us  s%d  %s\n", ms * 1e3, marks[i].stream, marks[i].name);
        }
        for (auto& m : marks) cudaEventDestroy(m.ev);
      }
    } else if (ctx->msm_graph_exec && memcmp(&key, &ctx->msm_key, sizeof(key)) == 0) {
      ZC_CUDA(ctx, cudaGraphLaunch((cudaGraphExec_t)ctx->msm_graph_exec, st));
      ctx->launches += ctx->msm_graph_launches;
    } else {
      ZC_CUDA(ctx, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      int32_t rc = enqueue();
      cudaGraph_t graph = nullptr;
      cudaError_t e = cudaStreamEndCapture(st, &graph);
      if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
      ZC_CUDA(ctx, e);
      // New arguments, same shape (new scalars / points / output buffer -- the common case): retarget the instantiated graph
      // in place.  cudaGraphExecUpdate touches no device allocation; destroying and re-instantiating an executable graph
      // frees device memory, which synchronises the DEVICE -- a stall per call, and a deadlock (until the exchange's deadline)
      // when another rank of the same process is already spinning in its exchange kernel.
      cudaGraphExec_t exec = (cudaGraphExec_t)ctx->msm_graph_exec;
      bool updated = false;
      if (exec && ctx->msm_key.n == key.n && ctx->msm_key.c == key.c && ctx->msm_key.rank == key.rank && ctx->msm_key.nranks == key.nranks &&
          ctx->msm_key.mode == key.mode && ctx->msm_key.ws == key.ws) {
        cudaGraphExecUpdateResultInfo info;
        if (cudaGraphExecUpdate(exec, graph, &info) == cudaSuccess) updated = true;
        else cudaGetLastError();                                // topology differs after all: instantiate below
      }
      if (!updated) {
        if (exec) { cudaGraphExecDestroy(exec); ctx->msm_graph_exec = nullptr; }
        exec = nullptr;
        e = cudaGraphInstantiate(&exec, graph, 0);
        if (e != cudaSuccess) { cudaGraphDestroy(graph); ZC_CUDA(ctx, e); }
        ctx->msm_graph_exec = exec;
      }
      cudaGraphDestroy(graph);
      ctx->msm_key = key;
      ctx->msm_graph_launches = nlaunch;
      ZC_CUDA(ctx, cudaGraphLaunch(exec, st));
      ctx->launches += nlaunch;
    }
    if (trace_mode == 2 && stamp_buf && !stamp_names.empty()) {
      cudaStreamSynchronize(st);
      std::vector<unsigned long long> h(stamp_names.size());
      cudaMemcpy(h.data(), stamp_buf, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      fprintf(stderr, "[zc_msm graph timeline] n=%zu c=%d rank %d/%d\n", n, c, rank, nranks);
      for (size_t i = 1; i < h.size(); i++)
        fprintf(stderr, "  %8.1f us  s%d  %s\n", (double)(h[i] - h[0]) * 1e-3, stamp_names[i].second, stamp_names[i].first);
    }
  }

  if (exchange && ctx->peers_connected) {
    // peer stores over NVLink + flag wait + tree fold in one kernel (zc_peer.cu)
    int32_t rc = zc_peer_exchange_fold(ctx, partial, out_point_dev);
    if (rc) return rc;
  } else if (exchange) {
    int32_t rc = zc_nccl_allgather(ctx, partial, ctx->gather_buf, 160);
    if (rc) return rc;
    rc = zc_point_fold_dev(ctx, (const uint64_t*)ctx->gather_buf, (size_t)nranks, out_point_dev);
    if (rc) return rc;
  }
  return ZC_OK;
}

extern "C" {

// The plan rules as host functions (no device needed): which rank owns a window and how a short window is spread.  The
// Python / C++ mirrors and a caller that pre-extracts one rank's digits ask the library instead of restating the rules.
int32_t zc_msm_plan_query(int32_t window_bits, int32_t window, int32_t nranks, int32_t out[4]) {
  if (!out) return ZC_ERR_NULL;
  if (window_bits < 8 || window_bits > 16) return ZC_ERR_MODE;
  const int nwin = (256 + window_bits - 1) / window_bits;
  if (window < 0 || window >= nwin || nranks < 1) return ZC_ERR_SIZE;
  out[0] = window_owner(window, nranks);
  out[1] = short_window_sub_bits(window_bits, window);
  out[2] = merged_spread_bits(window_bits, window);
  out[3] = window_bits * window - merged_spread_bits(window_bits, window);      // a fixed-base table row of this window is 2^out[3] P_i
  return ZC_OK;
}

// ---- fixed generators: an opaque handle that owns everything derived from the points ---------------------------------
int32_t zc_msm_generators_create_dev(zc_ctx *ctx, const uint64_t *points, size_t n, int32_t kind, int32_t window_bits,
                                     int32_t rank, int32_t nranks, zc_msm_generators **out) {
  if (!ctx) return ZC_ERR_NULL;
  ZC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!out) return zc_fail(ctx, ZC_ERR_NULL, "null handle pointer");
  *out = nullptr;
  if (!points || n == 0) return zc_fail(ctx, ZC_ERR_NULL, "null / empty points");
  if (n > ((size_t)1 << 31) - 1) return zc_fail(ctx, ZC_ERR_SIZE, "n exceeds 2^31 - 1");
  if (kind != ZC_GEN_PREPARED && kind != ZC_GEN_FIXED_BASE) return zc_fail(ctx, ZC_ERR_MODE, "kind must be ZC_GEN_PREPARED or ZC_GEN_FIXED_BASE");
  const int c = window_bits;
  FbWindows win;
  win.nwl = 0;
  if (kind == ZC_GEN_FIXED_BASE) {
    if (c < 8 || c > 16) return zc_fail(ctx, ZC_ERR_MODE, "window_bits must be in 8..16");
    if (nranks < 1 || rank < 0 || rank >= nranks) return zc_fail(ctx, ZC_ERR_SIZE, "bad rank / nranks");
    const int nwin = (256 + c - 1) / c;
    for (int w = 0; w < nwin; w++) if (window_owner(w, nranks) == rank) win.w[win.nwl++] = (int16_t)w;
    for (int t = win.nwl; t < MAX_WINDOWS; t++) win.w[t] = 0;
    if ((size_t)win.nwl * n > ((size_t)1 << 31) - 1) return zc_fail(ctx, ZC_ERR_SIZE, "windows x points exceeds 2^31 - 1");
  }
  zc_msm_generators *g = new zc_msm_generators();
  g->id = ctx->gens_next_id++; g->owner = ctx; g->kind = kind; g->n = n;
  g->c = kind == ZC_GEN_FIXED_BASE ? c : 0; g->rank = kind == ZC_GEN_FIXED_BASE ? rank : 0; g->nranks = kind == ZC_GEN_FIXED_BASE ? nranks : 0;
  g->cached = nullptr; g->table = nullptr; g->table_bytes = 0; g->corr = nullptr;
  auto fail = [&](int32_t rc) { cudaFree(g->cached); cudaFree(g->table); cudaFree(g->corr); delete g; return rc; };
#define ZC_GEN_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); return fail(-(int32_t)e__); } } while (0)
  if (kind == ZC_GEN_PREPARED) {
    // normalise to Z = 1 once (one inversion per point): every later bucket addition is a 7-multiplication mixed addition
    ZC_GEN_CUDA(cudaMalloc(&g->cached, n * 128));
    msm_prep_affine_kernel<<<(unsigned)(((n + PREP_K - 1) / PREP_K + 127) / 128), 128, 0, ctx->stream>>>(points, (uint32_t*)g->cached, n);
    ctx->launches++;
    ZC_GEN_CUDA(cudaGetLastError());
  } else {
    g->table_bytes = (size_t)win.nwl * n * 128;
    if (win.nwl > 0) {
      ZC_GEN_CUDA(cudaMalloc(&g->table, g->table_bytes));
      msm_fixed_base_table_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(points, (uint32_t*)g->table, n, c, win);
      ctx->launches++;
      ZC_GEN_CUDA(cudaGetLastError());
    }
    // spread short window (at most one per scalar width): the constant it adds to every MSM, negated, kept for the chain kernel
    for (int t = 0; t < win.nwl; t++) {
      const int sm = merged_spread_bits(c, win.w[t]);
      if (sm == 0) continue;
      void *ks = nullptr;
      int32_t rc;
      if ((rc = zc_scratch(ctx, 1, n * 40 + 8, &ks))) return fail(rc);
      ZC_GEN_CUDA(cudaMalloc(&g->corr, 160));
      msm_spread_scalars_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((uint64_t*)ks, n, sm, c * win.w[t] - sm);
      ctx->launches++;
      if ((rc = zc_msm_run(ctx, nullptr, points, (const uint64_t*)ks, n, c, 0, 1, false, (uint64_t*)g->corr))) return fail(rc);
      if ((rc = zc_point_neg_batch_dev(ctx, (const uint64_t*)g->corr, (uint64_t*)g->corr, 1))) return fail(rc);
    }
  }
  // the caller's point array is not referenced after this returns
  ZC_GEN_CUDA(cudaStreamSynchronize(ctx->stream));
#undef ZC_GEN_CUDA
  ctx->gens_live++;
  *out = g;
  return ZC_OK;
}

int32_t zc_msm_generators_destroy(zc_ctx *ctx, zc_msm_generators *gens) {
  if (!ctx) return ZC_ERR_NULL;
  if (!gens) return ZC_OK;
  if (gens->owner != ctx) return zc_fail(ctx, ZC_ERR_STATE, "generator handle belongs to another context");
  ZC_CUDA(ctx, cudaSetDevice(ctx->device));
  ZC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));            // an enqueued MSM may still read the tables
  if (ctx->msm_graph_exec && ctx->msm_key.gens_id == gens->id) { cudaGraphExecDestroy((cudaGraphExec_t)ctx->msm_graph_exec); ctx->msm_graph_exec = nullptr; }
  cudaFree(gens->cached); cudaFree(gens->table); cudaFree(gens->corr);
  gens->owner = nullptr;
  delete gens;
  ctx->gens_live--;
  return ZC_OK;
}

int32_t zc_msm_generators_info(const zc_msm_generators *gens, size_t *n, int32_t *kind, size_t *device_bytes) {
  if (!gens) return ZC_ERR_NULL;
  if (n) *n = gens->n;
  if (kind) *kind = gens->kind;
  if (device_bytes) *device_bytes = (gens->cached ? gens->n * 128 : 0) + gens->table_bytes + (gens->corr ? 160 : 0);
  return ZC_OK;
}

int32_t zc_msm_gen_dev(zc_ctx *ctx, const zc_msm_generators *gens, const uint64_t *scalars, int32_t window_bits, uint64_t *out_point_dev) {
  if (!ctx) return ZC_ERR_NULL;
  ZC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!gens || !scalars || !out_point_dev) return zc_fail(ctx, ZC_ERR_NULL, "null pointer argument");
  return zc_msm_run(ctx, gens, nullptr, scalars, gens->n, window_bits, 0, 1, false, out_point_dev);
}

int32_t zc_msm_gen_partial_dev(zc_ctx *ctx, const zc_msm_generators *gens, const uint64_t *scalars, int32_t window_bits,
                               int32_t rank, int32_t nranks, uint64_t *out_point_dev) {
  if (!ctx) return ZC_ERR_NULL;
  ZC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!gens || !scalars || !out_point_dev) return zc_fail(ctx, ZC_ERR_NULL, "null pointer argument");
  return zc_msm_run(ctx, gens, nullptr, scalars, gens->n, window_bits, rank, nranks, false, out_point_dev);
}

int32_t zc_msm_gen_sharded_dev(zc_ctx *ctx, const zc_msm_generators *gens, const uint64_t *scalars, int32_t window_bits, uint64_t *out_point_dev) {
  if (!ctx) return ZC_ERR_NULL;
  ZC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!gens || !scalars || !out_point_dev) return zc_fail(ctx, ZC_ERR_NULL, "null pointer argument");
  return zc_msm_run(ctx, gens, nullptr, scalars, gens->n, window_bits, ctx->rank, ctx->nranks, true, out_point_dev);
}

int32_t zc_msm_dev(zc_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, int32_t window_bits, uint64_t *out_point_dev) {
  if (!ctx) return ZC_ERR_NULL;
  ZC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!out_point_dev || (n && (!points || !scalars))) return zc_fail(ctx, ZC_ERR_NULL, "null pointer argument");
  return zc_msm_run(ctx, nullptr, points, scalars, n, window_bits, 0, 1, false, out_point_dev);
}

int32_t zc_msm_partial_dev(zc_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, int32_t window_bits,
                           int32_t rank, int32_t nranks, uint64_t *out_point_dev) {
  if (!ctx) return ZC_ERR_NULL;
  ZC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!out_point_dev || (n && (!points || !scalars))) return zc_fail(ctx, ZC_ERR_NULL, "null pointer argument");
  return zc_msm_run(ctx, nullptr, points, scalars, n, window_bits, rank, nranks, false, out_point_dev);
}

int32_t zc_msm_sharded_dev(zc_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, int32_t window_bits, uint64_t *out_point_dev) {
  if (!ctx) return ZC_ERR_NULL;
  ZC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!out_point_dev || (n && (!points || !scalars))) return zc_fail(ctx, ZC_ERR_NULL, "null pointer argument");
  return zc_msm_run(ctx, nullptr, points, scalars, n, window_bits, ctx->rank, ctx->nranks, true, out_point_dev);
}

// host-pointer MSM: the scalars travel first so that digit extraction and the sort run under the (4x larger) copy of the
// points; then one run on the resident arrays.  sharded != 0: every rank passes the same arrays (collective).
static int32_t msm_host(zc_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, int32_t window_bits, bool sharded, uint64_t *out_point) {
  if (!ctx) return ZC_ERR_NULL;
  ZC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!out_point || (n && (!points || !scalars))) return zc_fail(ctx, ZC_ERR_NULL, "null pointer argument");
  void *dp = nullptr, *ds = nullptr, *dout = nullptr;
  int32_t rc;
  if ((rc = zc_scratch(ctx, 0, n * 160 + 8, &dp))) return rc;
  if ((rc = zc_scratch(ctx, 1, n * 40 + 8, &ds))) return rc;
  if ((rc = zc_scratch(ctx, 2, 160, &dout))) return rc;
  if (n) {
    ZC_CUDA(ctx, cudaMemcpyAsync(ds, scalars, n * 40, cudaMemcpyHostToDevice, ctx->stream));
    ZC_CUDA(ctx, cudaMemcpyAsync(dp, points, n * 160, cudaMemcpyHostToDevice, ctx->stream));
  }
  if ((rc = zc_msm_run(ctx, nullptr, (const uint64_t*)dp, (const uint64_t*)ds, n, window_bits, sharded ? ctx->rank : 0, sharded ? ctx->nranks : 1,
                       sharded, (uint64_t*)dout))) return rc;
  ZC_CUDA(ctx, cudaMemcpyAsync(out_point, dout, 160, cudaMemcpyDeviceToHost, ctx->stream));
  ZC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return sharded ? zc_peer_check_error(ctx) : ZC_OK;
}

int32_t zc_msm(zc_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, int32_t window_bits, uint64_t *out_point) {
  return msm_host(ctx, points, scalars, n, window_bits, false, out_point);
}
int32_t zc_msm_sharded(zc_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, int32_t window_bits, uint64_t *out_point) {
  return msm_host(ctx, points, scalars, n, window_bits, true, out_point);
}

}  // extern "C"
