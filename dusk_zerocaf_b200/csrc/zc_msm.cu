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
constexpr int CHUNK_LOG = 3;         // m = 8 buckets per chunk in the reduce levels
constexpr int CHUNK = 1 << CHUNK_LOG;

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
    v = pt_add_fast(v, o);
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
      if (!have) { acc = v; have = true; } else acc = pt_add_fast(acc, v);
    }
    acc = warp_sum_pt(acc);
    if (lane == 0) st_pt(buckets + 32 * (size_t)g, acc);
  }
}

// ---- reduce level: chunks of CHUNK items -> (sum, weighted-from-zero sum) ----------------------------------------
// For chunk items I_0..I_{m-1}:  sums = sum I_j,  acc0 = sum j * I_j   (descending running sum)
__global__ void __launch_bounds__(128) msm_chunk_kernel(const uint32_t* __restrict__ in, int n_in, int nwl,
                                                        uint32_t* __restrict__ sums, uint32_t* __restrict__ acc0) {
  const int n_out = (n_in + CHUNK - 1) / CHUNK;
  size_t g = (size_t)blockIdx.x * 128 + threadIdx.x;
  if (g >= (size_t)nwl * n_out) return;
  size_t wl = g / n_out, t = g - wl * n_out;
  const uint32_t* base = in + 32 * (wl * n_in);
  int lo = (int)t * CHUNK;
  int hi = min(lo + CHUNK, n_in);
  Pt run = ld_pt(base + 32 * (size_t)(hi - 1));
  Pt acc = run;
  if (hi - 1 == lo) acc = pt_identity_mont();
  for (int k = hi - 2; k >= lo; k--) {
    run = pt_add_fast(run, ld_pt(base + 32 * (size_t)k));
    if (k > lo) acc = pt_add_fast(acc, run);
  }
  // acc = sum_{k>lo} (k-lo) I_k  (for hi-1 > lo the loop added run for k = hi-2 .. lo+1 on top of the initial I_{hi-1})
  st_pt(sums + 32 * g, run);
  st_pt(acc0 + 32 * g, acc);
}

// ---- plain sum of n items per window (one block per window) ---------------------------------------------------------
constexpr int SUM_TPB = 256;
__global__ void __launch_bounds__(SUM_TPB) msm_sum_kernel(const uint32_t* __restrict__ in, int n_in, uint32_t* __restrict__ out, int out_stride, int out_slot) {
  __shared__ uint32_t sm[SUM_TPB * 32];
  const int wl = blockIdx.x;
  const uint32_t* base = in + 32 * ((size_t)wl * n_in);
  Pt acc = pt_identity_mont();
  bool have = false;
  for (int k = threadIdx.x; k < n_in; k += SUM_TPB) {
    Pt v = ld_pt(base + 32 * (size_t)k);
    if (!have) { acc = v; have = true; } else acc = pt_add_fast(acc, v);
  }
  st_pt(sm + 32 * threadIdx.x, acc);
  __syncthreads();
  for (int d = SUM_TPB / 2; d >= 1; d >>= 1) {
    if (threadIdx.x < d) {
      Pt a = ld_pt(sm + 32 * threadIdx.x);
      Pt b = ld_pt(sm + 32 * (threadIdx.x + d));
      st_pt(sm + 32 * threadIdx.x, pt_add_fast(a, b));
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) st_pt(out + 32 * ((size_t)wl * out_stride + out_slot), ld_pt(sm));
}

// ---- combine: per window Horner over the levels, then Horner over the windows with 2^c scalings -------------------
// level_sums[wl][l] (l = 0..nlev-1) = S_l = sum_t acc0^(l)_t ;  level_sums[wl][nlev] = Total = sum of all buckets.
// window sum  W = Total + sum_l m^l S_l.   partial = sum_{local w} 2^(c w) W_w.
__global__ void msm_combine_kernel(const uint32_t* __restrict__ level_sums, int nlev, int nwl, int c, int nwin,
                                   int rank, int nranks, uint64_t* __restrict__ out52, uint32_t* __restrict__ wsum) {
  const int wl = threadIdx.x;
  if (wl < nwl) {
    const uint32_t* ls = level_sums + 32 * ((size_t)wl * (nlev + 1));
    Pt r = ld_pt(ls + 32 * (size_t)(nlev - 1));
    for (int l = nlev - 2; l >= 0; l--) {
      for (int j = 0; j < CHUNK_LOG; j++) r = pt_double_fast(r);
      r = pt_add_fast(r, ld_pt(ls + 32 * (size_t)l));
    }
    r = pt_add_fast(r, ld_pt(ls + 32 * (size_t)nlev));
    st_pt(wsum + 32 * wl, r);
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  Pt acc = pt_identity_mont();
  bool started = false;
  for (int w = nwin - 1; w >= 0; w--) {
    if (started) for (int j = 0; j < c; j++) acc = pt_double_fast(acc);
    if (w % nranks == rank) {
      Pt ww = ld_pt(wsum + 32 * (size_t)(w / nranks));
      if (started) acc = pt_add_fast(acc, ww); else { acc = ww; started = true; }
    }
  }
  // Montgomery-form words read as normal form are the same point scaled by R: store them directly.
  pt_store52(out52, acc);
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

  // level geometry
  int lev_n[16]; int nlev = 0;
  { int cur = nb; while (cur > 1) { cur = (cur + CHUNK - 1) / CHUNK; lev_n[nlev++] = cur; } }

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
    size_t o_heavy = o; o = align_up(o + 256 + (size_t)nwl * nb * 4, 256);
    size_t o_hist = o;   o = align_up(o + (size_t)nwl * nb * 4, 256);
    size_t o_offs = o;   o = align_up(o + (size_t)nwl * nb * 4, 256);
    size_t o_cursor = o; o = align_up(o + (size_t)nwl * nb * 4, 256);
    size_t o_buckets = o; o = align_up(o + (size_t)nwl * nb * 128, 256);
    size_t o_sums[2]; size_t o_acc0;
    o_sums[0] = o; o = align_up(o + (size_t)nwl * lev_n[0] * 128, 256);
    o_sums[1] = o; o = align_up(o + (size_t)nwl * lev_n[0] * 128, 256);
    o_acc0 = o;    o = align_up(o + (size_t)nwl * lev_n[0] * 128, 256);
    size_t o_lsum = o; o = align_up(o + (size_t)nwl * (nlev + 1) * 128, 256);
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
    uint32_t *heavy_count = (uint32_t*)(ws + o_heavy);
    uint32_t *heavy_list = heavy_count + 64;
    uint32_t *offs = (uint32_t*)(ws + o_offs);
    uint32_t *cursor = (uint32_t*)(ws + o_cursor);
    uint32_t *buckets = (uint32_t*)(ws + o_buckets);
    uint32_t *sums[2] = {(uint32_t*)(ws + o_sums[0]), (uint32_t*)(ws + o_sums[1])};
    uint32_t *acc0 = (uint32_t*)(ws + o_acc0);
    uint32_t *lsum = (uint32_t*)(ws + o_lsum);
    uint32_t *wsum = (uint32_t*)(ws + o_wsum);
    cudaStream_t st = ctx->stream;

    ZC_CUDA(ctx, cudaMemsetAsync(hist, 0, (size_t)nwl * nb * 4, st));
    ZC_CUDA(ctx, cudaMemsetAsync(heavy_count, 0, 256, st));
    msm_prep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(points, cached, n); ctx->launches++;
    msm_digits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(scalars, n, c, nwin, rank, nranks, digits, hist); ctx->launches++;
    msm_scan_kernel<<<nwl, 1024, 0, st>>>(hist, offs, cursor, nb); ctx->launches++;
    {
      size_t tot = n * (size_t)nwl;
      msm_scatter_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(digits, n, n_pad, nwl, nb, cursor, sorted); ctx->launches++;
    }
    {
      size_t tot = (size_t)nwl * nb;
      size_t tseg = (size_t)nwl * nseg;
      msm_accum_kernel<<<(unsigned)((tseg + ACC_TPB - 1) / ACC_TPB), ACC_TPB, 0, st>>>(cached, sorted, offs, hist, n_pad, nseg, nwl, nb, buckets, partH, partT); ctx->launches++;
      msm_fix_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(offs, hist, nseg, nwl, nb, partH, partT, buckets, heavy_count, heavy_list); ctx->launches++;
      msm_heavy_kernel<<<2 * ctx->sm_count, 128, 0, st>>>(offs, hist, nseg, nb, partH, partT, buckets, heavy_count, heavy_list); ctx->launches++;
    }
    // reduce levels
    const uint32_t *in = buckets; int n_in = nb;
    for (int l = 0; l < nlev; l++) {
      int n_out = lev_n[l];
      size_t tot = (size_t)nwl * n_out;
      msm_chunk_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(in, n_in, nwl, sums[l & 1], acc0); ctx->launches++;
      msm_sum_kernel<<<nwl, SUM_TPB, 0, st>>>(acc0, n_out, lsum, nlev + 1, l); ctx->launches++;
      in = sums[l & 1]; n_in = n_out;
    }
    // Total = the single item left at the top level
    msm_sum_kernel<<<nwl, SUM_TPB, 0, st>>>(in, n_in, lsum, nlev + 1, nlev); ctx->launches++;
    msm_combine_kernel<<<1, 32, 0, st>>>(lsum, nlev, nwl, c, nwin, rank, nranks, partial, wsum); ctx->launches++;
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
