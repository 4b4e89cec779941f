// zc_msm.cu -- Pippenger multi-scalar multiplication  sum_i [s_i] P_i  on one GPU, or one rank's share of the
// windows when the MSM is sharded by bucket-window over several GPUs.
//
// The reference has no MSM: the semantics are fold(Add, identity, [double_and_add(P_i, s_i)])
// (/root/reference/src/edwards.rs:102-120, 465-489) as a group element.  Fast formulas are used throughout (cached-operand
// addition, dedicated doubling), so the result is compared canonically (affine / Ristretto equality), not limb-wise.
//
// Pipeline (all on ctx->stream):
//   prep     points (AoS radix-2^52, normal form) -> cached operands (Y+X, Y-X, Z, 2dT) as 32 x u32.  A normal-form
//            coordinate vector IS a Montgomery-form representation of the same projective point (all four coordinates
//            scaled by 1/R), so no conversion multiply is needed -- only the 2d*T product.
//   digits   scalars -> signed c-bit digits for this rank's windows (+ per-bucket histogram, global atomics)
//   scan     exclusive scan of the histogram per window (bucket start offsets)
//   scatter  counting-sort scatter of (point index | sign) into bucket order
//   accum    one thread per 32-entry segment of the sorted list (balanced), 8M cached additions, prefetched gathers;
//            fix/heavy stitch buckets that span segments
//   reduce   sum_k (k+1) * B_k per window: multi-level chunked running sums (chunk m), tree sums per level
//   combine  Horner over the levels and over this rank's windows with 2^c scalings -> one partial point
//   exchange (sharded only) ncclAllGather of the partial points + fixed-order fold with the reference Add
#include "zc_internal.h"
#include "zc_point.cuh"

using namespace zc;

int32_t zc_nccl_allgather(zc_ctx *ctx, const void *send, void *recv, size_t bytes);   // zc_nccl.cu

namespace {

constexpr int MAX_WINDOWS = 32;      // ceil(256 / 8)
constexpr int MAX_GROUPS = 4;        // window groups processed top-down; the scaling chain of one group overlaps the next

struct PtW { uint32_t w[32]; };      // packed point: 4 coordinates x 8 words (extended or cached, Montgomery form)

__device__ __forceinline__ void ld_fe(const uint32_t* __restrict__ p, Fe& a) {
  uint4 lo = *reinterpret_cast<const uint4*>(p);
  uint4 hi = *reinterpret_cast<const uint4*>(p + 4);
  a.w[0] = lo.x; a.w[1] = lo.y; a.w[2] = lo.z; a.w[3] = lo.w;
  a.w[4] = hi.x; a.w[5] = hi.y; a.w[6] = hi.z; a.w[7] = hi.w;
}
__device__ __forceinline__ void st_fe(uint32_t* __restrict__ p, const Fe& a) {
  *reinterpret_cast<uint4*>(p) = make_uint4(a.w[0], a.w[1], a.w[2], a.w[3]);
  *reinterpret_cast<uint4*>(p + 4) = make_uint4(a.w[4], a.w[5], a.w[6], a.w[7]);
}
__device__ __forceinline__ Pt ld_pt(const uint32_t* __restrict__ p) {
  Pt r; ld_fe(p, r.X); ld_fe(p + 8, r.Y); ld_fe(p + 16, r.Z); ld_fe(p + 24, r.T); return r;
}
__device__ __forceinline__ void st_pt(uint32_t* __restrict__ p, const Pt& a) {
  st_fe(p, a.X); st_fe(p + 8, a.Y); st_fe(p + 16, a.Z); st_fe(p + 24, a.T);
}
__device__ __forceinline__ PtCached ld_cached(const uint32_t* __restrict__ p) {
  PtCached r; ld_fe(p, r.YpX); ld_fe(p + 8, r.YmX); ld_fe(p + 16, r.Z); ld_fe(p + 24, r.T2d); return r;
}

// 1/d * R mod p: recovers 2T from the cached 2dT when a bucket is initialised from its first point
__device__ __forceinline__ Fe DINV_MONT() {
  return Fe{{0x69c50bb0u, 0xa53327e2u, 0x96b47422u, 0xeaa0ffd5u, 0xfd35fb8fu, 0xd34f1e03u, 0x8d35344bu, 0x0b7245f4u}};
}
// the point a cached operand stands for, as (2X, 2Y, 2Z, 2T)
__device__ __forceinline__ Pt cached_to_pt(const PtCached& c) {
  typedef ModP M;
  Pt r;
  r.X = fe_sub<M>(c.YpX, c.YmX);
  r.Y = fe_add<M>(c.YpX, c.YmX);
  r.Z = fe_add<M>(c.Z, c.Z);
  r.T = mont_mul<M>(c.T2d, DINV_MONT());
  return r;
}

// ---- prep ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) msm_prep_kernel(const uint64_t* __restrict__ points, uint32_t* __restrict__ cached, size_t n) {
  size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  Pt p = pt_load52(points + 20 * i);
  PtCached c = pt_to_cached(p);
  uint32_t* o = cached + 32 * i;
  st_fe(o, c.YpX); st_fe(o + 8, c.YmX); st_fe(o + 16, c.Z); st_fe(o + 24, c.T2d);
}

// ---- digits + histogram ----------------------------------------------------------------------------------------
// digit d_w in [-2^(c-1), 2^(c-1)):  s = sum_w d_w 2^(c w).  Bucket slot = |d| - 1 in [0, 2^(c-1)).
__global__ void __launch_bounds__(256) msm_digits_kernel(const uint64_t* __restrict__ scalars, size_t n, int c, int nwin,
                                                         int rank, int nranks, int32_t* __restrict__ digits,
                                                         uint32_t* __restrict__ hist) {
  size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  Fe s = fe_load52(scalars + 5 * i);
  const uint32_t half = 1u << (c - 1);
  const uint32_t mask = (1u << c) - 1u;
  const uint32_t nb = half;
  uint32_t carry = 0;
  int wl = 0;
  for (int w = 0; w < nwin; w++) {
    int bit = w * c;
    int word = bit >> 5, sh = bit & 31;
    uint32_t raw = 0;
    if (word < 8) {
      uint64_t two = s.w[word];
      if (word + 1 < 8) two |= (uint64_t)s.w[word + 1] << 32;
      raw = (uint32_t)(two >> sh) & mask;
    }
    raw += carry;
    int32_t d;
    if (raw >= half) { d = (int32_t)raw - (int32_t)(1u << c); carry = 1; } else { d = (int32_t)raw; carry = 0; }
    if (w % nranks == rank) {
      digits[(size_t)wl * n + i] = d;
      if (d != 0) {
        uint32_t slot = (uint32_t)(d < 0 ? -d : d) - 1u;
        atomicAdd(&hist[(size_t)wl * nb + slot], 1u);
      }
      wl++;
    }
  }
}

// ---- exclusive scan of each window's histogram (one block of 1024 threads per local window) -----------------------
__global__ void __launch_bounds__(1024) msm_scan_kernel(const uint32_t* __restrict__ hist, uint32_t* __restrict__ offs,
                                                        uint32_t* __restrict__ cursor, int nb) {
  __shared__ uint32_t part[1024];
  const int wl = blockIdx.x;
  const uint32_t* h = hist + (size_t)wl * nb;
  const int per = (nb + 1023) / 1024;
  const int lo = threadIdx.x * per;
  uint32_t sum = 0;
  for (int k = 0; k < per; k++) if (lo + k < nb) sum += h[lo + k];
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    uint32_t v = (threadIdx.x >= d) ? part[threadIdx.x - d] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = part[threadIdx.x] - sum;
  for (int k = 0; k < per; k++) {
    if (lo + k < nb) {
      offs[(size_t)wl * nb + lo + k] = run;
      cursor[(size_t)wl * nb + lo + k] = run;
      run += h[lo + k];
    }
  }
}

// ---- scatter: counting sort of point indices into bucket order ----------------------------------------------------
__global__ void __launch_bounds__(256) msm_scatter_kernel(const int32_t* __restrict__ digits, size_t n, size_t n_pad, int nwl, int nb,
                                                          uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted) {
  size_t g = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (g >= n * (size_t)nwl) return;
  size_t wl = g / n, i = g - wl * n;
  int32_t d = digits[g];
  if (d == 0) return;
  uint32_t slot = (uint32_t)(d < 0 ? -d : d) - 1u;
  uint32_t pos = atomicAdd(&cursor[wl * nb + slot], 1u);
  sorted[wl * n_pad + pos] = (uint32_t)i | (d < 0 ? 0x80000000u : 0u);
}

// ---- bucket accumulation, balanced: one thread per SEGMENT of SEG consecutive sorted entries ------------------------
// A thread walks its SEG entries, summing runs of equal bucket.  A run that covers its whole bucket is stored straight
// into buckets[]; a run cut by a segment boundary goes to the segment's H slot (run containing the segment's first
// entry) or T slot (run containing its last entry, when that is a different run).  msm_fix_kernel then stitches the
// buckets that span several segments.  Every thread does the same number of additions, whatever the digit
// distribution (a top window with few, heavy buckets used to serialise thousands of additions in one thread).
constexpr int SEG = 32;
constexpr int ACC_TPB = 128;

__device__ __forceinline__ PtCached ld_entry(const uint32_t* __restrict__ cached, uint32_t e) {
  PtCached c = ld_cached(cached + 32 * (size_t)(e & 0x7fffffffu));
  if (e >> 31) c = pt_cached_neg(c);
  return c;
}

__global__ void __launch_bounds__(ACC_TPB) msm_accum_kernel(const uint32_t* __restrict__ cached, const uint32_t* __restrict__ sorted,
                                                            const uint32_t* __restrict__ offs, const uint32_t* __restrict__ hist,
                                                            size_t n_pad, int nseg, int nwl, int nb,
                                                            uint32_t* __restrict__ buckets, uint32_t* __restrict__ partH,
                                                            uint32_t* __restrict__ partT) {
  __shared__ uint32_t idx_s[SEG * ACC_TPB];
  const int tx = threadIdx.x;
  size_t g = (size_t)blockIdx.x * ACC_TPB + tx;
  if (g >= (size_t)nwl * nseg) return;
  const size_t wl = g / nseg;
  const uint32_t s = (uint32_t)(g - wl * nseg);
  const uint32_t* woffs = offs + wl * nb;
  const uint32_t* whist = hist + wl * nb;
  const uint32_t nnz = woffs[nb - 1] + whist[nb - 1];
  const uint32_t start = s * SEG;
  if (start >= nnz) return;
  const uint32_t end = min(start + SEG, nnz);
  {
    const uint4* src = reinterpret_cast<const uint4*>(sorted + wl * n_pad + start);
#pragma unroll
    for (int j = 0; j < SEG / 4; j++) {
      uint4 v = src[j];
      idx_s[(4 * j + 0) * ACC_TPB + tx] = v.x; idx_s[(4 * j + 1) * ACC_TPB + tx] = v.y;
      idx_s[(4 * j + 2) * ACC_TPB + tx] = v.z; idx_s[(4 * j + 3) * ACC_TPB + tx] = v.w;
    }
  }
  // bucket of the first entry: last b with offs[b] <= start  (upper_bound - 1)
  uint32_t lo = 0, hi = (uint32_t)nb;
  while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (woffs[mid] <= start) lo = mid + 1; else hi = mid; }
  uint32_t b = lo - 1;
  uint32_t bbeg = woffs[b], bend = bbeg + whist[b];
  uint32_t run_start = start;
  uint32_t* bk = buckets + 32 * (wl * nb);
  uint32_t* pH = partH + 32 * g;
  uint32_t* pT = partT + 32 * g;

  PtCached cur = ld_entry(cached, idx_s[tx]);
  Pt acc = cached_to_pt(cur);
  if (start + 1 < end) cur = ld_entry(cached, idx_s[ACC_TPB + tx]);
#pragma unroll 1
  for (uint32_t k = start + 1; k < end; k++) {
    PtCached nxt = cur;
    if (k + 1 < end) nxt = ld_entry(cached, idx_s[(k + 1 - start) * ACC_TPB + tx]);   // prefetch: independent of acc
    if (k == bend) {
      // flush the finished run
      const bool complete = (run_start == bbeg);
      st_pt(complete ? bk + 32 * (size_t)b : pH, acc);   // incomplete here <=> the run began before this segment (H slot)
      do { b++; } while (whist[b] == 0);
      bbeg = k; bend = k + whist[b];
      run_start = k;
      acc = cached_to_pt(cur);
    } else {
      acc = pt_add_cached(acc, cur);
    }
    cur = nxt;
  }
  {
    const bool complete = (run_start == bbeg) && (end == bend);
    uint32_t* dst = complete ? bk + 32 * (size_t)b : (run_start == start ? pH : pT);
    st_pt(dst, acc);
  }
}

// ---- stitch buckets that span several segments; write the identity into empty buckets -----------------------------
// bucket range [o, e): s_first = o / SEG, s_last = (e-1) / SEG.  If s_first == s_last the run was complete and is already
// in buckets[].  Otherwise  sum = (o == s_first*SEG ? H[s_first] : T[s_first]) + H[s_first+1] + ... + H[s_last].
// Buckets with more than FIX_INLINE partials are queued for msm_heavy_kernel (one warp per bucket, tree sum).
constexpr int FIX_INLINE = 6;
__global__ void __launch_bounds__(128) msm_fix_kernel(const uint32_t* __restrict__ offs, const uint32_t* __restrict__ hist,
                                                      int nseg, int nwl, int nb, const uint32_t* __restrict__ partH,
                                                      const uint32_t* __restrict__ partT, uint32_t* __restrict__ buckets,
                                                      uint32_t* __restrict__ heavy_count, uint32_t* __restrict__ heavy_list) {
  size_t g = (size_t)blockIdx.x * 128 + threadIdx.x;
  if (g >= (size_t)nwl * nb) return;
  const size_t wl = g / nb;
  const uint32_t cnt = hist[g];
  if (cnt == 0) { st_pt(buckets + 32 * g, pt_identity_mont()); return; }
  const uint32_t o = offs[g], e = o + cnt;
  const uint32_t s_first = o / SEG, s_last = (e - 1) / SEG;
  if (s_first == s_last) return;
  if (s_last - s_first + 1 > FIX_INLINE) {
    uint32_t slot = atomicAdd(heavy_count, 1u);
    heavy_list[slot] = (uint32_t)g;
    return;
  }
  const uint32_t* H = partH + 32 * (wl * nseg);
  const uint32_t* T = partT + 32 * (wl * nseg);
  Pt acc = ld_pt((o == s_first * SEG ? H : T) + 32 * (size_t)s_first);
  for (uint32_t s = s_first + 1; s <= s_last; s++) acc = pt_add_fast(acc, ld_pt(H + 32 * (size_t)s));
  st_pt(buckets + 32 * g, acc);
}

// Out-of-line point operations for the latency-bound tail kernels (reduce / heavy): one copy of the ~30 KB addition
// body per kernel keeps them inside the instruction cache (inlined at every call site, msm_reduce1_kernel was 1 MB of
// straight-line code and ran at ~8 cycles per instruction).
__device__ __noinline__ Pt pt_add_ni(Pt p, Pt q) { return pt_add_fast(p, q); }
__device__ __noinline__ Pt pt_double_ni(Pt p) { return pt_double_fast(p); }

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

__global__ void __launch_bounds__(128) msm_heavy_kernel(const uint32_t* __restrict__ offs, const uint32_t* __restrict__ hist,
                                                        int nseg, int nb, const uint32_t* __restrict__ partH,
                                                        const uint32_t* __restrict__ partT, uint32_t* __restrict__ buckets,
                                                        const uint32_t* __restrict__ heavy_count, const uint32_t* __restrict__ heavy_list) {
  const uint32_t nheavy = *heavy_count;
  const int lane = threadIdx.x & 31;
  for (uint32_t h = blockIdx.x * 4 + (threadIdx.x >> 5); h < nheavy; h += gridDim.x * 4) {
    const uint32_t g = heavy_list[h];
    const size_t wl = g / (uint32_t)nb;
    const uint32_t o = offs[g], e = o + hist[g];
    const uint32_t s_first = o / SEG, s_last = (e - 1) / SEG;
    const uint32_t* H = partH + 32 * (wl * nseg);
    const uint32_t* T = partT + 32 * (wl * nseg);
    Pt acc = pt_identity_mont();
    bool have = false;
    for (uint32_t s = s_first + lane; s <= s_last; s += 32) {
      Pt v = ld_pt(((s == s_first && o != s_first * SEG) ? T : H) + 32 * (size_t)s);
      if (!have) { acc = v; have = true; } else acc = pt_add_ni(acc, v);
    }
    acc = warp_sum_pt(acc);
    if (lane == 0) st_pt(buckets + 32 * (size_t)g, acc);
  }
}

// lane-to-lane copies of a field element / point
__device__ __forceinline__ Fe shfl_fe(const Fe& a, int src) {
  Fe r;
#pragma unroll
  for (int k = 0; k < 8; k++) r.w[k] = __shfl_sync(0xffffffffu, a.w[k], src);
  return r;
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

// ---- bucket reduction  W = sum_k (k+1) B_k,  two warp-cooperative levels ------------------------------------------
// A warp owns RCH = 32 * RQ consecutive items.  Lane l walks its RQ items with the running-sum trick
//   c_l = sum_j I_j,   a_l = sum_j j I_j          (2 RQ - 1 additions, serial)
// then the warp takes a suffix scan S_l = sum_{i >= l} c_i (5 shuffle steps) and one tree sum:
//   sum_k (k - base) I_k = sum_l (a_l + RQ * l * c_l) = sum_l a_l + RQ * sum_{l >= 1} S_l.
// Depth per level ~ 2 RQ + 14 additions instead of the 5 chunk + 5 tree launches this replaces.
constexpr int RQ = 8;
constexpr int RCH = 32 * RQ;      // 256 buckets per warp at level 0

template <int Q, bool WITH_PLAIN>
__device__ __forceinline__ void warp_weighted(const uint32_t* __restrict__ items, const uint32_t* __restrict__ plain,
                                              int n_items, int lane, Pt& total, Pt& weighted, Pt& plain_sum) {
  // lane's items: indices lane*Q .. lane*Q+Q-1  (identity beyond n_items)
  const int lo = lane * Q;
  Pt run = pt_identity_mont(), acc = pt_identity_mont(), pl = pt_identity_mont();
#pragma unroll 1
  for (int j = Q - 1; j >= 0; j--) {
    const int k = lo + j;
    if (k < n_items) {
      run = pt_add_ni(run, ld_pt(items + 32 * (size_t)k));
      if (WITH_PLAIN) pl = pt_add_ni(pl, ld_pt(plain + 32 * (size_t)k));
    }
    if (j > 0) acc = pt_add_ni(acc, run);
  }
  // suffix scan of run over lanes
  Pt S = run;
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) {
    Pt o = shfl_down_pt(S, d);
    Pt t = pt_add_ni(S, o);
    if (lane + d < 32) S = t;
  }
  // z = Q * (lane >= 1 ? S : 0) + acc
  Pt z = (lane >= 1) ? S : pt_identity_mont();
#pragma unroll
  for (int q = Q; q > 1; q >>= 1) z = pt_double_ni(z);
  z = pt_add_ni(z, acc);
  weighted = warp_sum_pt(z);
  total = S;                 // valid on lane 0
  if (WITH_PLAIN) plain_sum = warp_sum_pt(pl);
}

// level 0: one warp per chunk of RCH buckets -> chunk sum C_t and chunk-local weighted sum Wt_t = sum (k - base) B_k
__global__ void __launch_bounds__(128) msm_reduce0_kernel(const uint32_t* __restrict__ buckets, int nb, int nchunk, int nwl,
                                                          uint32_t* __restrict__ csum, uint32_t* __restrict__ wloc) {
  const int lane = threadIdx.x & 31;
  const size_t g = (size_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (g >= (size_t)nwl * nchunk) return;
  const size_t wl = g / nchunk;
  const int t = (int)(g - wl * nchunk);
  const int base = t * RCH;
  Pt total, weighted, unused;
  warp_weighted<RQ, false>(buckets + 32 * (wl * nb + base), nullptr, min(RCH, nb - base), lane, total, weighted, unused);
  if (lane == 0) { st_pt(csum + 32 * g, total); st_pt(wloc + 32 * g, weighted); }
}

// level 1: one warp per window.  W = sum_t [ Wt_t + (t RCH + 1) C_t ] = sum_t Wt_t + sum_t C_t + RCH * sum_t t C_t
template <int Q1>
__device__ __forceinline__ Pt reduce1_body(const uint32_t* __restrict__ csum, const uint32_t* __restrict__ wloc, int nchunk, int lane) {
  Pt total, weighted, plain;
  warp_weighted<Q1, true>(csum, wloc, nchunk, lane, total, weighted, plain);
#pragma unroll 1
  for (int q = RCH; q > 1; q >>= 1) weighted = pt_double_ni(weighted);
  return pt_add_ni(pt_add_ni(weighted, total), plain);   // valid on lane 0
}
__global__ void __launch_bounds__(32) msm_reduce1_kernel(const uint32_t* __restrict__ csum, const uint32_t* __restrict__ wloc,
                                                         int nchunk, uint32_t* __restrict__ wsum) {
  const int lane = threadIdx.x;
  const size_t wl = blockIdx.x;
  const uint32_t* c = csum + 32 * (wl * nchunk);
  const uint32_t* w = wloc + 32 * (wl * nchunk);
  Pt r;
  if (nchunk <= 32) r = reduce1_body<1>(c, w, nchunk, lane);
  else if (nchunk <= 64) r = reduce1_body<2>(c, w, nchunk, lane);
  else r = reduce1_body<4>(c, w, nchunk, lane);
  if (lane == 0) st_pt(wsum + 32 * wl, r);
}

// ---- window chain: acc = 2^(c w) - weighted sum of this rank's window sums, four lanes per point operation ----------
// Lane q = lane & 3 holds coordinate q (X, Y, Z, T) of the running point; the four independent field multiplications
// of each of the two stages of a doubling / addition run on the four lanes, the operands travel by shuffle.  This
// turns the strictly serial tail (c doublings per window) from ~8 dependent multiplications per operation into 2.
__device__ __forceinline__ Fe quad_stage2(const Fe& E, const Fe& F, const Fe& G, const Fe& H, int q) {
  typedef ModP M;
  // X3 = E F, Y3 = G H, Z3 = F G, T3 = E H
  Fe u, v;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    u.w[k] = (q == 0 || q == 3) ? E.w[k] : (q == 1 ? G.w[k] : F.w[k]);
    v.w[k] = (q == 0) ? F.w[k] : (q == 2 ? G.w[k] : H.w[k]);
  }
  return mont_mul<M>(u, v);
}
__device__ __forceinline__ Fe quad_double(const Fe& c, int q, int qbase) {
  typedef ModP M;
  Fe x = shfl_fe(c, qbase), y = shfl_fe(c, qbase + 1);
  Fe in = c;
  if (q == 3) in = fe_add<M>(x, y);
  Fe s = mont_mul<M>(in, in);                                  // A, B, ZZ, (X+Y)^2 on lanes 0..3
  Fe A = shfl_fe(s, qbase), B = shfl_fe(s, qbase + 1), ZZ = shfl_fe(s, qbase + 2), SS = shfl_fe(s, qbase + 3);
  Fe C = fe_add<M>(ZZ, ZZ);
  Fe D = fe_neg<M>(A);
  Fe E = fe_sub<M>(fe_sub<M>(SS, A), B);
  Fe G = fe_add<M>(D, B);
  Fe F = fe_sub<M>(G, C);
  Fe H = fe_sub<M>(D, B);
  return quad_stage2(E, F, G, H, q);
}
// c (distributed over the quad) += the full point p (every lane holds all of p)
__device__ __forceinline__ Fe quad_add(const Fe& c, const Pt& p, int q, int qbase) {
  typedef ModP M;
  Fe x1 = shfl_fe(c, qbase), y1 = shfl_fe(c, qbase + 1);
  Fe u, v;
  if (q == 0)      { u = fe_sub<M>(y1, x1); v = fe_sub<M>(p.Y, p.X); }
  else if (q == 1) { u = fe_add<M>(y1, x1); v = fe_add<M>(p.Y, p.X); }
  else if (q == 2) { u = c;                 v = fe_add<M>(p.Z, p.Z); }
  else             { u = c;                 v = mont_mul<M>(p.T, D2_MONT()); }
  Fe s = mont_mul<M>(u, v);                                    // A, B, D = 2 Z1 Z2, C = T1 2d T2
  Fe A = shfl_fe(s, qbase), B = shfl_fe(s, qbase + 1), D = shfl_fe(s, qbase + 2), C = shfl_fe(s, qbase + 3);
  Fe E = fe_sub<M>(B, A);
  Fe F = fe_sub<M>(D, C);
  Fe G = fe_add<M>(D, C);
  Fe H = fe_add<M>(B, A);
  return quad_stage2(E, F, G, H, q);
}

// One warp.  acc (in/out, extended Montgomery words) is the running sum already scaled to this group's top window.
//   for i in 0..ng-1:  if (i > 0) acc = 2^gap_in acc;   acc += wsum[i]       (group windows in descending order)
//   acc = 2^gap_post acc
// first != 0: acc starts as the identity.  out52 != nullptr: also store the result in the ABI layout.
__global__ void __launch_bounds__(32) msm_chain_kernel(const uint32_t* __restrict__ wsum, int ng, int wsum_step, int first,
                                                       int gap_in, int gap_post, uint32_t* __restrict__ acc_io,
                                                       uint64_t* __restrict__ out52) {
  const int lane = threadIdx.x;
  const int q = lane & 3, qbase = lane & ~3;
  Pt a0 = first ? pt_identity_mont() : ld_pt(acc_io);
  Fe c = (q == 0) ? a0.X : (q == 1 ? a0.Y : (q == 2 ? a0.Z : a0.T));
#pragma unroll 1
  for (int i = 0; i < ng; i++) {
    if (i > 0) {
#pragma unroll 1
      for (int j = 0; j < gap_in; j++) c = quad_double(c, q, qbase);
    }
    Pt w = ld_pt(wsum + 32 * ((ptrdiff_t)i * wsum_step));
    c = quad_add(c, w, q, qbase);
  }
#pragma unroll 1
  for (int j = 0; j < gap_post; j++) c = quad_double(c, q, qbase);
  Pt r;
  r.X = shfl_fe(c, 0); r.Y = shfl_fe(c, 1); r.Z = shfl_fe(c, 2); r.T = shfl_fe(c, 3);
  if (lane == 0) {
    st_pt(acc_io, r);
    // Montgomery-form words read as normal form are the same point scaled by R: store them directly.
    if (out52) pt_store52(out52, r);
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int32_t zc_msm_run(zc_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, int32_t c,
                   int32_t rank, int32_t nranks, bool exchange, uint64_t *out_point_dev) {
  if (c < 8 || c > 16) return zc_fail(ctx, ZC_ERR_MODE, "window_bits must be in 8..16");
  if (nranks < 1 || rank < 0 || rank >= nranks) return zc_fail(ctx, ZC_ERR_SIZE, "bad rank / nranks");
  if (n > ((size_t)1 << 31) - 1) return zc_fail(ctx, ZC_ERR_SIZE, "n exceeds 2^31 - 1");
  const int nwin = (256 + c - 1) / c;
  const int nb = 1 << (c - 1);
  int nwl = 0;
  for (int w = 0; w < nwin; w++) if (w % nranks == rank) nwl++;

  uint64_t *partial = exchange ? nullptr : out_point_dev;
  if (exchange) {
    if (!ctx->nccl_comm) return zc_fail(ctx, ZC_ERR_STATE, "zc_msm_sharded_dev needs zc_ctx_set_nccl first");
    partial = (uint64_t*)ctx->gather_buf + 20 * (size_t)nranks;   // send slot after the nranks receive slots
  }

  if (nwl == 0 || n == 0) {
    // this rank owns no window (nranks > nwin) or the MSM is empty: contribute the identity
    int32_t rc = zc_point_fold_dev(ctx, nullptr, 0, partial);
    if (rc) return rc;
  } else {
    // workspace layout
    size_t o = 0;
    size_t o_cached = o; o = align_up(o + n * 128, 256);
    size_t o_digits = o; o = align_up(o + (size_t)nwl * n * 4, 256);
    const size_t n_pad = align_up(n, SEG);
    const int nseg = (int)(n_pad / SEG);
    size_t o_sorted = o; o = align_up(o + (size_t)nwl * n_pad * 4 + 256, 256);
    size_t o_partH = o; o = align_up(o + (size_t)nwl * nseg * 128, 256);
    size_t o_partT = o; o = align_up(o + (size_t)nwl * nseg * 128, 256);
    size_t o_heavy = o; o = align_up(o + 256 * MAX_GROUPS + (size_t)nwl * nb * 4, 256);
    size_t o_hist = o;   o = align_up(o + (size_t)nwl * nb * 4, 256);
    size_t o_offs = o;   o = align_up(o + (size_t)nwl * nb * 4, 256);
    size_t o_cursor = o; o = align_up(o + (size_t)nwl * nb * 4, 256);
    size_t o_buckets = o; o = align_up(o + (size_t)nwl * nb * 128, 256);
    const int nchunk = (nb + RCH - 1) / RCH;
    size_t o_csum = o; o = align_up(o + (size_t)nwl * nchunk * 128, 256);
    size_t o_wloc = o; o = align_up(o + (size_t)nwl * nchunk * 128, 256);
    size_t o_acc = o;  o = align_up(o + 128, 256);
    size_t o_wsum = o; o = align_up(o + (size_t)nwl * 128, 256);
    if (o > ctx->msm_ws_bytes) {
      if (ctx->msm_ws) ZC_CUDA(ctx, cudaFree(ctx->msm_ws));
      ctx->msm_ws = nullptr; ctx->msm_ws_bytes = 0;
      ZC_CUDA(ctx, cudaMalloc(&ctx->msm_ws, o));
      ctx->msm_ws_bytes = o;
    }
    char *ws = (char*)ctx->msm_ws;
    uint32_t *cached = (uint32_t*)(ws + o_cached);
    int32_t *digits = (int32_t*)(ws + o_digits);
    uint32_t *sorted = (uint32_t*)(ws + o_sorted);
    uint32_t *hist = (uint32_t*)(ws + o_hist);
    uint32_t *partH = (uint32_t*)(ws + o_partH);
    uint32_t *partT = (uint32_t*)(ws + o_partT);
    uint32_t *heavy_count = (uint32_t*)(ws + o_heavy);          // one counter per group, 256 B apart
    uint32_t *heavy_list = heavy_count + 64 * MAX_GROUPS;
    uint32_t *offs = (uint32_t*)(ws + o_offs);
    uint32_t *cursor = (uint32_t*)(ws + o_cursor);
    uint32_t *buckets = (uint32_t*)(ws + o_buckets);
    uint32_t *csum = (uint32_t*)(ws + o_csum);
    uint32_t *wloc = (uint32_t*)(ws + o_wloc);
    uint32_t *acc = (uint32_t*)(ws + o_acc);
    uint32_t *wsum = (uint32_t*)(ws + o_wsum);
    cudaStream_t st = ctx->stream;
    if (!ctx->side_stream) {
      ZC_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
      for (int i = 0; i < MAX_GROUPS + 1; i++) ZC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev[i], cudaEventDisableTiming));
    }
    cudaStream_t side = ctx->side_stream;

    ZC_CUDA(ctx, cudaMemsetAsync(hist, 0, (size_t)nwl * nb * 4, st));
    ZC_CUDA(ctx, cudaMemsetAsync(heavy_count, 0, 256 * MAX_GROUPS, st));
    msm_prep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(points, cached, n); ctx->launches++;
    msm_digits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(scalars, n, c, nwin, rank, nranks, digits, hist); ctx->launches++;
    msm_scan_kernel<<<nwl, 1024, 0, st>>>(hist, offs, cursor, nb); ctx->launches++;
    {
      size_t tot = n * (size_t)nwl;
      msm_scatter_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(digits, n, n_pad, nwl, nb, cursor, sorted); ctx->launches++;
    }
    // Window groups, top-down.  Local window wl is global window rank + nranks * wl.  After a group's buckets are
    // accumulated the side stream stitches and reduces them, folds the window sums into acc and scales acc down to the
    // next group's top window (or, after the last group, by 2^(c * rank)) while the main stream accumulates the next group.
    const int ngroups = nwl < MAX_GROUPS ? nwl : MAX_GROUPS;
    int hi = nwl;
    for (int g = 0; g < ngroups; g++) {
      const int gsz = (hi + (ngroups - g) - 1) / (ngroups - g);
      const int lo = hi - gsz;
      const size_t tot = (size_t)gsz * nb;
      const size_t tseg = (size_t)gsz * nseg;
      const uint32_t *g_sorted = sorted + (size_t)lo * n_pad, *g_offs = offs + (size_t)lo * nb, *g_hist = hist + (size_t)lo * nb;
      uint32_t *g_buckets = buckets + 32 * ((size_t)lo * nb), *g_partH = partH + 32 * ((size_t)lo * nseg), *g_partT = partT + 32 * ((size_t)lo * nseg);
      uint32_t *g_hcount = heavy_count + 64 * g, *g_hlist = heavy_list + (size_t)lo * nb;
      msm_accum_kernel<<<(unsigned)((tseg + ACC_TPB - 1) / ACC_TPB), ACC_TPB, 0, st>>>(cached, g_sorted, g_offs, g_hist, n_pad, nseg, gsz, nb, g_buckets, g_partH, g_partT); ctx->launches++;
      // everything after the accumulation is latency-bound (few warps, long dependent chains): it runs on the side
      // stream, under the next group's accumulation
      ZC_CUDA(ctx, cudaEventRecord(ctx->ev[g], st));
      ZC_CUDA(ctx, cudaStreamWaitEvent(side, ctx->ev[g], 0));
      msm_fix_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, side>>>(g_offs, g_hist, nseg, gsz, nb, g_partH, g_partT, g_buckets, g_hcount, g_hlist); ctx->launches++;
      msm_heavy_kernel<<<2 * ctx->sm_count, 128, 0, side>>>(g_offs, g_hist, nseg, nb, g_partH, g_partT, g_buckets, g_hcount, g_hlist); ctx->launches++;
      msm_reduce0_kernel<<<(unsigned)(((size_t)gsz * nchunk + 3) / 4), 128, 0, side>>>(g_buckets, nb, nchunk, gsz, csum + 32 * ((size_t)lo * nchunk), wloc + 32 * ((size_t)lo * nchunk)); ctx->launches++;
      msm_reduce1_kernel<<<gsz, 32, 0, side>>>(csum + 32 * ((size_t)lo * nchunk), wloc + 32 * ((size_t)lo * nchunk), nchunk, wsum + 32 * (size_t)lo); ctx->launches++;
      const bool last = (g == ngroups - 1);
      msm_chain_kernel<<<1, 32, 0, side>>>(wsum + 32 * (size_t)(hi - 1), gsz, -1, g == 0 ? 1 : 0, c * nranks, last ? c * rank : c * nranks,
                                           acc, last ? partial : nullptr); ctx->launches++;
      hi = lo;
    }
    ZC_CUDA(ctx, cudaEventRecord(ctx->ev[MAX_GROUPS], side));
    ZC_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev[MAX_GROUPS], 0));
    ZC_CUDA(ctx, cudaGetLastError());
  }

  if (exchange) {
    int32_t rc = zc_nccl_allgather(ctx, partial, ctx->gather_buf, 160);
    if (rc) return rc;
    rc = zc_point_fold_dev(ctx, (const uint64_t*)ctx->gather_buf, (size_t)nranks, out_point_dev);
    if (rc) return rc;
  }
  return ZC_OK;
}

extern "C" {

int32_t zc_msm_dev(zc_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, int32_t window_bits, uint64_t *out_point_dev) {
  if (!ctx) return ZC_ERR_NULL;
  ZC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!out_point_dev || (n && (!points || !scalars))) return zc_fail(ctx, ZC_ERR_NULL, "null pointer argument");
  return zc_msm_run(ctx, points, scalars, n, window_bits, 0, 1, false, out_point_dev);
}

int32_t zc_msm_partial_dev(zc_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, int32_t window_bits,
                           int32_t rank, int32_t nranks, uint64_t *out_point_dev) {
  if (!ctx) return ZC_ERR_NULL;
  ZC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!out_point_dev || (n && (!points || !scalars))) return zc_fail(ctx, ZC_ERR_NULL, "null pointer argument");
  return zc_msm_run(ctx, points, scalars, n, window_bits, rank, nranks, false, out_point_dev);
}

int32_t zc_msm_sharded_dev(zc_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, int32_t window_bits, uint64_t *out_point_dev) {
  if (!ctx) return ZC_ERR_NULL;
  ZC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!out_point_dev || (n && (!points || !scalars))) return zc_fail(ctx, ZC_ERR_NULL, "null pointer argument");
  return zc_msm_run(ctx, points, scalars, n, window_bits, ctx->rank, ctx->nranks, true, out_point_dev);
}

int32_t zc_msm(zc_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, int32_t window_bits, uint64_t *out_point) {
  if (!ctx) return ZC_ERR_NULL;
  ZC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!out_point || (n && (!points || !scalars))) return zc_fail(ctx, ZC_ERR_NULL, "null pointer argument");
  void *dp = nullptr, *ds = nullptr, *dout = nullptr;
  int32_t rc;
  if ((rc = zc_scratch(ctx, 0, n * 160 + 8, &dp))) return rc;
  if ((rc = zc_scratch(ctx, 1, n * 40 + 8, &ds))) return rc;
  if ((rc = zc_scratch(ctx, 2, 160, &dout))) return rc;
  if (n) {
    ZC_CUDA(ctx, cudaMemcpyAsync(dp, points, n * 160, cudaMemcpyHostToDevice, ctx->stream));
    ZC_CUDA(ctx, cudaMemcpyAsync(ds, scalars, n * 40, cudaMemcpyHostToDevice, ctx->stream));
  }
  if ((rc = zc_msm_run(ctx, (const uint64_t*)dp, (const uint64_t*)ds, n, window_bits, 0, 1, false, (uint64_t*)dout))) return rc;
  ZC_CUDA(ctx, cudaMemcpyAsync(out_point, dout, 160, cudaMemcpyDeviceToHost, ctx->stream));
  ZC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZC_OK;
}

}  // extern "C"
