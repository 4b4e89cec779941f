// zc_peer.cu -- the exchange step of the bucket-window-sharded MSM over NVLink peer memory, fused with the fold.
//
// Every rank owns a small mailbox in its own HBM, mapped into all peers with CUDA IPC.  One warp per rank
//   1. stores its 160-byte partial point into slot[rank] of EVERY rank's mailbox (peer stores over NVLink / NVSwitch),
//   2. publishes flag[rank] = sequence number on every rank (after a system-scope fence),
//   3. waits until all flags of its own mailbox carry the sequence number,
//   4. folds the partial points in a fixed tree with the reference Add (edwards.rs:465-489) -- the same tree on every
//      rank, so all ranks return identical bits.
// One kernel, no host round trip, no library collective: for a 160-byte payload the cost is NVLink latency, not bandwidth.
// zc_msm_sharded_dev takes this path when the mailboxes are connected and falls back to ncclAllGather + fold otherwise.
#include "zc_internal.h"
#include "zc_point.cuh"

using namespace zc;

namespace {

__device__ __forceinline__ void st_sys_u64(uint64_t* p, uint64_t v) { asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ uint64_t ld_sys_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(32) msm_exchange_fold_kernel(zc_peer_ptrs peers, int rank, int nranks,
                                                               const uint64_t* __restrict__ partial, uint64_t* __restrict__ out) {
  const int lane = threadIdx.x;
  zc_mailbox* mine = peers.p[rank];
  // sequence number of this exchange: every rank calls the collective the same number of times
  uint64_t seq = 0;
  if (lane == 0) { seq = mine->counter + 1; mine->counter = seq; }
  seq = __shfl_sync(0xffffffffu, seq, 0);
  // 1. my partial into slot[rank] of every mailbox (lanes 0..19 carry the 20 limbs)
  if (lane < 20) {
    const uint64_t v = partial[lane];
    for (int r = 0; r < nranks; r++) st_sys_u64(&peers.p[r]->slot[rank][lane], v);
  }
  __threadfence_system();
  __syncwarp();
  // 2. publish
  if (lane < nranks) st_sys_u64(&peers.p[lane]->flag[rank], seq);
  // 3. wait for everybody's partial to land here
  if (lane < nranks) { while (ld_sys_u64(&mine->flag[lane]) < seq) { } }
  __syncwarp();
  __threadfence_system();
  // 4. fixed-shape tree fold: lane r starts with partial r (identity beyond nranks), log2 steps of the reference Add
  Pt acc = pt_identity_mont();
  if (lane < nranks) {
    uint64_t l[20];
#pragma unroll
    for (int k = 0; k < 20; k++) l[k] = ld_sys_u64(&mine->slot[lane][k]);
    acc = pt_to_mont(Pt{fe_from_limbs52(l[0], l[1], l[2], l[3], l[4]), fe_from_limbs52(l[5], l[6], l[7], l[8], l[9]),
                        fe_from_limbs52(l[10], l[11], l[12], l[13], l[14]), fe_from_limbs52(l[15], l[16], l[17], l[18], l[19])});
  }
  int width = 1;
  while (width < nranks) width <<= 1;
#pragma unroll 1
  for (int d = width >> 1; d >= 1; d >>= 1) {
    Pt o;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      o.X.w[k] = __shfl_down_sync(0xffffffffu, acc.X.w[k], d);
      o.Y.w[k] = __shfl_down_sync(0xffffffffu, acc.Y.w[k], d);
      o.Z.w[k] = __shfl_down_sync(0xffffffffu, acc.Z.w[k], d);
      o.T.w[k] = __shfl_down_sync(0xffffffffu, acc.T.w[k], d);
    }
    acc = pt_add_ref(acc, o);
  }
  if (lane == 0) pt_store52(out, pt_from_mont(acc));
}

}  // namespace

int32_t zc_peer_exchange_fold(zc_ctx* ctx, const uint64_t* partial, uint64_t* out) {
  msm_exchange_fold_kernel<<<1, 32, 0, ctx->stream>>>(ctx->peers, ctx->rank, ctx->nranks, partial, out);
  ctx->launches++;
  ZC_CUDA(ctx, cudaGetLastError());
  return ZC_OK;
}

extern "C" {

int32_t zc_peer_mailbox_create(zc_ctx* ctx, uint8_t handle_out[64]) {
  if (!ctx || !handle_out) return ZC_ERR_NULL;
  ZC_CUDA(ctx, cudaSetDevice(ctx->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!ctx->mailbox) {
    ZC_CUDA(ctx, cudaMalloc(&ctx->mailbox, sizeof(zc_mailbox)));
    ZC_CUDA(ctx, cudaMemset(ctx->mailbox, 0, sizeof(zc_mailbox)));
  }
  cudaIpcMemHandle_t h;
  ZC_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->mailbox));
  memcpy(handle_out, &h, 64);
  return ZC_OK;
}

int32_t zc_peer_mailbox_connect(zc_ctx* ctx, const uint8_t* handles, int32_t rank, int32_t nranks) {
  if (!ctx || !handles) return ZC_ERR_NULL;
  if (nranks < 1 || nranks > ZC_MAX_PEERS || rank < 0 || rank >= nranks) return zc_fail(ctx, ZC_ERR_SIZE, "bad rank / nranks (at most 16 peers)");
  if (!ctx->mailbox) return zc_fail(ctx, ZC_ERR_STATE, "zc_peer_mailbox_create first");
  ZC_CUDA(ctx, cudaSetDevice(ctx->device));
  for (int r = 0; r < nranks; r++) {
    if (r == rank) { ctx->peers.p[r] = (zc_mailbox*)ctx->mailbox; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * (size_t)r, 64);
    void* p = nullptr;
    ZC_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->peers.p[r] = (zc_mailbox*)p;
  }
  if (!ctx->gather_buf) ZC_CUDA(ctx, cudaMalloc(&ctx->gather_buf, (size_t)(nranks + 1) * 160));
  ctx->rank = rank;
  ctx->nranks = nranks;
  ctx->peers_connected = true;
  return ZC_OK;
}

}  // extern "C"
