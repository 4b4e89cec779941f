// zc_peer.cu -- the exchange step of the bucket-window-sharded MSM over NVLink peer memory, fused with the fold.
//
// Every rank owns a small mailbox in its own HBM, mapped into all peers with CUDA IPC.  One warp per rank
//   1. stores its 160-byte partial point into slot[rank] of EVERY rank's mailbox (peer stores over NVLink / NVSwitch),
//   2. publishes flag[rank] = sequence number on every rank (after a system-scope fence),
//   3. waits until all flags of its own mailbox carry the sequence number (bounded: a missing rank becomes an error, not a hang),
//   4. folds the partial points in a fixed tree (the group law of edwards.rs:465-489, four lanes per point) -- the same
//      operations in the same order on every rank, so all ranks return identical bits.
// One kernel, no host round trip, no library collective: for a 160-byte payload the cost is NVLink latency, not bandwidth.
// zc_msm_sharded_dev takes this path when the mailboxes are connected and falls back to ncclAllGather + fold otherwise.
#include <stdlib.h>

#include "zc_internal.h"
#include "zc_quad.cuh"

using namespace zc;

namespace {

__device__ __forceinline__ void st_sys_u64(uint64_t* p, uint64_t v) { asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void st_release_sys_u64(uint64_t* p, uint64_t v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ uint64_t ld_sys_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t ld_acquire_sys_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long peer_timer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Slots and flags are double-buffered by the parity of the sequence number: a rank that has finished exchange k may run
// ahead into exchange k+1 and store into slot[(k+1) & 1] while a slower peer still reads slot[k & 1]; it cannot get to
// k+2 (same parity as k) before that peer has published its flag for k+1, i.e. has left exchange k.  The wait is bounded:
// after timeout_ns without a flag the kernel records (sequence, missing rank) in the context's error word (mapped host
// memory, checked by zc_ctx_sync and the host-pointer entry points) and returns the identity instead of hanging.
__global__ void __launch_bounds__(32) msm_exchange_fold_kernel(zc_peer_ptrs peers, int rank, int nranks, uint64_t seq, uint64_t timeout_ns,
                                                               const uint64_t* __restrict__ partial, uint64_t* __restrict__ out,
                                                               uint64_t* __restrict__ error_word) {
  const int lane = threadIdx.x, q = lane & 3, qbase = lane & 28, j = lane >> 2;
  const int par = (int)(seq & 1);
  zc_mailbox* mine = peers.p[rank];
  // 1. my partial into slot[par][rank] of every mailbox (lanes 0..19 carry the 20 limbs)
  if (lane < 20) {
    const uint64_t v = partial[lane];
    for (int r = 0; r < nranks; r++) st_sys_u64(&peers.p[r]->slot[par][rank][lane], v);
  }
  __threadfence_system();
  __syncwarp();
  // 2. publish
  if (lane < nranks) st_release_sys_u64(&peers.p[lane]->flag[par][rank], seq);
  // 3. wait for everybody's partial to land here, with a deadline
  int missing = 0;
  if (lane < nranks) {
    const unsigned long long t0 = peer_timer();
    while (ld_acquire_sys_u64(&mine->flag[par][lane]) < seq) {
      if (peer_timer() - t0 > timeout_ns) { missing = lane + 1; break; }
    }
  }
  const unsigned late = __ballot_sync(0xffffffffu, missing != 0);
  __threadfence_system();
  if (late) {
    if (lane == 0) {
      *error_word = (seq << 8) | (uint64_t)__ffs(late);          // first missing rank + 1
      __threadfence_system();
      for (int k = 0; k < 20; k++) out[k] = (k == 5 || k == 10) ? 1ull : 0ull;     // identity (0, 1, 1, 0)
    }
    return;
  }
  // 4. fixed-shape fold, four lanes per point: quad j starts with partial j (+ partial j + 8), then a tree over the quads.
  //    The same operations in the same order on every rank: identical bits everywhere.  Normal-form limbs are a
  //    Montgomery-form representative of the same projective point, so no conversion is needed in or out.
  auto load_slot = [&](int r) {
    uint64_t l[20];
#pragma unroll
    for (int k = 0; k < 20; k++) l[k] = ld_sys_u64(&mine->slot[par][r][k]);
    return Pt{fe_from_limbs52(l[0], l[1], l[2], l[3], l[4]), fe_from_limbs52(l[5], l[6], l[7], l[8], l[9]),
              fe_from_limbs52(l[10], l[11], l[12], l[13], l[14]), fe_from_limbs52(l[15], l[16], l[17], l[18], l[19])};
  };
  const Fe zero{{0, 0, 0, 0, 0, 0, 0, 0}};
  const Fe ident = (q == 1 || q == 2) ? Consts<ModP>::R1() : zero;
  Fe acc = ident;
  if (j < nranks) {
    const Pt p = load_slot(j);
    acc = pt_coord(p, q);
  }
  if (nranks > 8) {                                                // warp-uniform
    const bool on = j + 8 < nranks;
    const Pt p = load_slot(on ? j + 8 : 0);
    const Fe r = quad_add(acc, p, q, qbase);
    if (on) acc = r;
  }
#pragma unroll 1
  for (int d = 4; d >= 1; d >>= 1) {
    if (d >= nranks) continue;                                     // warp-uniform: nothing above this level
    Pt o;
    const int src = ((j + d) & 7) * 4;
    o.X = shfl_fe(acc, src); o.Y = shfl_fe(acc, src + 1); o.Z = shfl_fe(acc, src + 2); o.T = shfl_fe(acc, src + 3);
    const Fe r = quad_add(acc, o, q, qbase);
    if (j < d) acc = r;
  }
  acc = fe_canon4(acc);
  Pt res;
  res.X = shfl_fe(acc, 0); res.Y = shfl_fe(acc, 1); res.Z = shfl_fe(acc, 2); res.T = shfl_fe(acc, 3);
  if (lane == 0) pt_store52(out, res);
}

}  // namespace

int32_t zc_peer_exchange_fold(zc_ctx* ctx, const uint64_t* partial, uint64_t* out) {
  static const uint64_t timeout_ms = getenv("ZC_PEER_TIMEOUT_MS") ? strtoull(getenv("ZC_PEER_TIMEOUT_MS"), nullptr, 10) : 5000ull;
  msm_exchange_fold_kernel<<<1, 32, 0, ctx->stream>>>(ctx->peers, ctx->rank, ctx->nranks, ctx->peer_seq, timeout_ms * 1000000ull,
                                                      partial, out, ctx->peer_error_dev);
  ctx->launches++;
  ZC_CUDA(ctx, cudaGetLastError());
  return ZC_OK;
}

// a timed-out exchange leaves (sequence << 8 | missing rank + 1) in the error word; reported once, then cleared
int32_t zc_peer_check_error(zc_ctx* ctx) {
  if (!ctx->peer_error_host) return ZC_OK;
  const uint64_t e = *(volatile uint64_t*)ctx->peer_error_host;
  if (e == 0) return ZC_OK;
  *(volatile uint64_t*)ctx->peer_error_host = 0;
  snprintf(ctx->err, sizeof(ctx->err), "sharded MSM exchange %llu timed out waiting for rank %d (result replaced by the identity)",
           (unsigned long long)(e >> 8), (int)(e & 0xff) - 1);
  return ZC_ERR_STATE;
}

extern "C" {

int32_t zc_peer_mailbox_create(zc_ctx* ctx, uint8_t handle_out[64]) {
  if (!ctx || !handle_out) return ZC_ERR_NULL;
  ZC_CUDA(ctx, cudaSetDevice(ctx->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  // Always a FRESH, zeroed mailbox and sequence numbers from 0: (re)connecting is create -> exchange handles -> connect on
  // every rank, and a mailbox is zeroed before its handle exists, so no rank can deliver into it too early.
  ZC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->peers_connected && ctx->peers_ipc)
    for (int r = 0; r < ctx->nranks; r++) if (r != ctx->rank && ctx->peers.p[r]) { cudaIpcCloseMemHandle(ctx->peers.p[r]); ctx->peers.p[r] = nullptr; }
  ctx->peers_connected = false;
  if (ctx->mailbox) { ZC_CUDA(ctx, cudaFree(ctx->mailbox)); ctx->mailbox = nullptr; }
  ZC_CUDA(ctx, cudaMalloc(&ctx->mailbox, sizeof(zc_mailbox)));
  ZC_CUDA(ctx, cudaMemset(ctx->mailbox, 0, sizeof(zc_mailbox)));
  ctx->peer_seq = 0;
  cudaIpcMemHandle_t h;
  ZC_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->mailbox));
  memcpy(handle_out, &h, 64);
  return ZC_OK;
}

}  // extern "C"

// ptrs != nullptr: the peers' mailboxes as plain device pointers (same process); else CUDA IPC handles (one process per GPU)
static int32_t peer_connect(zc_ctx* ctx, const uint8_t* handles, void* const* ptrs, int32_t rank, int32_t nranks) {
  if (!ctx || (!handles && !ptrs)) return ZC_ERR_NULL;
  if (nranks < 1 || nranks > ZC_MAX_PEERS || rank < 0 || rank >= nranks) return zc_fail(ctx, ZC_ERR_SIZE, "bad rank / nranks (at most 16 peers)");
  if (!ctx->mailbox) return zc_fail(ctx, ZC_ERR_STATE, "zc_peer_mailbox_create first");
  ZC_CUDA(ctx, cudaSetDevice(ctx->device));
  for (int r = 0; r < nranks; r++) {
    if (r == rank) { ctx->peers.p[r] = (zc_mailbox*)ctx->mailbox; continue; }
    if (ptrs) {
      if (!ptrs[r]) return zc_fail(ctx, ZC_ERR_NULL, "null peer mailbox");
      cudaPointerAttributes at;
      ZC_CUDA(ctx, cudaPointerGetAttributes(&at, ptrs[r]));
      if (at.type != cudaMemoryTypeDevice) return zc_fail(ctx, ZC_ERR_STATE, "peer mailbox is not device memory");
      if (at.device != ctx->device) {                               // another GPU driven by this process: map it
        cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError(); else ZC_CUDA(ctx, e);
      }
      ctx->peers.p[r] = (zc_mailbox*)ptrs[r];
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * (size_t)r, 64);
    void* p = nullptr;
    ZC_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->peers.p[r] = (zc_mailbox*)p;
  }
  ctx->peers_ipc = ptrs == nullptr;
  // always sized for THIS nranks (zc_ctx_set_nccl may have run earlier with fewer ranks)
  if (ctx->gather_buf) { ZC_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); ZC_CUDA(ctx, cudaFree(ctx->gather_buf)); ctx->gather_buf = nullptr; }
  ZC_CUDA(ctx, cudaMalloc(&ctx->gather_buf, (size_t)(nranks + 1) * 160));
  if (!ctx->peer_error_host) {
    ZC_CUDA(ctx, cudaHostAlloc(&ctx->peer_error_host, 64, cudaHostAllocMapped));
    memset(ctx->peer_error_host, 0, 64);
    ZC_CUDA(ctx, cudaHostGetDevicePointer((void**)&ctx->peer_error_dev, ctx->peer_error_host, 0));
  }
  ctx->rank = rank;
  ctx->nranks = nranks;
  ctx->peers_connected = true;
  return ZC_OK;
}

extern "C" {

int32_t zc_peer_mailbox_connect(zc_ctx* ctx, const uint8_t* handles, int32_t rank, int32_t nranks) {
  return peer_connect(ctx, handles, nullptr, rank, nranks);
}
int32_t zc_peer_mailbox_ptr(zc_ctx* ctx, void** out) {
  if (!ctx || !out) return ZC_ERR_NULL;
  if (!ctx->mailbox) return zc_fail(ctx, ZC_ERR_STATE, "zc_peer_mailbox_create first");
  *out = ctx->mailbox;
  return ZC_OK;
}
int32_t zc_peer_mailbox_connect_local(zc_ctx* ctx, void* const* mailboxes, int32_t rank, int32_t nranks) {
  return peer_connect(ctx, nullptr, mailboxes, rank, nranks);
}

}  // extern "C"
