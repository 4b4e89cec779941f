// zc_internal.h -- context object and launch helpers shared by the translation units of libzerocaf_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/zerocaf_b200.h"

#define ZC_PIPE_MAX_CHUNKS 512

// arguments an instantiated MSM graph was recorded with (zc_msm.cu)
struct zc_msm_key {
  const void *points, *scalars;
  size_t n;
  int32_t c, rank, nranks, mode;
  const void *partial, *ws;
  uint64_t gens_id;                   // 0 = no generator handle
};

// Fixed generators prepared once (zc_msm_generators_create_dev).  The handle OWNS everything derived from the points -- the
// caller's point array is not referenced after creation, so freeing, reusing or overwriting it cannot alias a later MSM.
struct zc_msm_generators {
  uint64_t id;                        // unique per creation (a recycled heap address never matches a recorded graph)
  zc_ctx *owner;
  int32_t kind, c, rank, nranks;      // c / rank / nranks: the call shape fixed-base tables were built for
  size_t n;
  void *cached;                       // n x 128 B affine cached operands (y+x, y-x, 1, 2dxy), Montgomery words
  void *table;                        // ZC_GEN_FIXED_BASE: rows 2^(c w) P_i for the windows of (rank, nranks), same record
  size_t table_bytes;
  void *corr;                         // [u64;20]: minus the constant a spread short window adds (zc_msm.cu), or null
};

// NVLink peer-memory exchange (zc_peer.cu): one mailbox per rank, mapped into every peer with CUDA IPC
#define ZC_MAX_PEERS 16
struct zc_mailbox {                     // double-buffered by the parity of the exchange's sequence number
  uint64_t flag[2][ZC_MAX_PEERS];     // flag[par][r] = sequence number of the last partial rank r delivered into slot[par][r]
  uint64_t slot[2][ZC_MAX_PEERS][20]; // slot[par][r] = rank r's partial point, [u64;20]
};
struct zc_peer_ptrs { zc_mailbox* p[ZC_MAX_PEERS]; };

struct zc_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  uint64_t launches = 0;
  char err[512] = {0};
  // grow-only device scratch for the host-pointer entry points (three inputs/outputs + MSM workspace)
  void *scratch[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t scratch_bytes[6] = {0, 0, 0, 0, 0, 0};
  // MSM workspace (device)
  void *msm_ws = nullptr;
  size_t msm_ws_bytes = 0;
  // NCCL (resolved with dlopen, see zc_nccl.cu)
  void *nccl_comm = nullptr;
  int rank = 0, nranks = 1;
  void *gather_buf = nullptr;   // nranks * 20 u64, device
  void *mailbox = nullptr;      // zc_mailbox in this rank's HBM
  zc_peer_ptrs peers = {};
  bool peers_connected = false;
  bool peers_ipc = false;             // peers.p[] were opened with cudaIpcOpenMemHandle (else plain pointers of this process)
  uint64_t peer_seq = 0;              // exchanges started on this context (every rank counts the same calls, failed ones included)
  void *peer_error_host = nullptr;    // mapped host word: (sequence << 8 | missing rank + 1) of a timed-out exchange
  uint64_t *peer_error_dev = nullptr;
  // host-pointer entry points: copy-in / copy-out streams and per-chunk events (created on first use)
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  cudaEvent_t pipe_ev[2 * ZC_PIPE_MAX_CHUNKS] = {};
  // input validation (zc_ctx_set_validation): first offending element index, ~0 = none
  bool validate = false, in_pipeline = false;
  size_t vbase = 0;
  unsigned long long *vflag_dev = nullptr;
  uint64_t gens_next_id = 1;        // ids of zc_msm_generators handles created on this context
  int gens_live = 0;
  void *basepoint_table = nullptr;  // fixed-base table (zc_fixed.cu), built on first use
  void *msm_graph_exec = nullptr;   // cudaGraphExec_t of the last MSM configuration
  zc_msm_key msm_key = {};
  uint64_t msm_graph_launches = 0;
  // MSM: side stream for the window-scaling chain + events (created on first use)
  cudaStream_t side_stream = nullptr, side_extra[3] = {nullptr, nullptr, nullptr}, chain_stream = nullptr, sort_stream = nullptr;
  cudaStream_t side_hi[3] = {nullptr, nullptr, nullptr};   // high-priority twins of the first three side streams (sharded MSM)
  cudaStream_t acc2 = nullptr;                            // second accumulation stream (ZC_MSM_ACC_OVERLAP: consecutive groups' accumulations back to back)
  cudaStream_t sort_hi = nullptr;                         // high-priority sort stream (sharded MSM: the first sort must not queue behind the operand pass)
  cudaEvent_t ev[16] = {};
};

#define ZC_CUDA(ctx, call)                                                                       \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess) {                                                                    \
      snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d %s -> %s", __FILE__, __LINE__, #call,      \
               cudaGetErrorString(e__));                                                         \
      return -(int32_t)e__;                                                                      \
    }                                                                                            \
  } while (0)

static inline int32_t zc_fail(zc_ctx *ctx, int32_t code, const char *msg) {
  if (ctx) snprintf(ctx->err, sizeof(ctx->err), "%s", msg);
  return code;
}

// grow-only scratch slot
static inline int32_t zc_scratch(zc_ctx *ctx, int slot, size_t bytes, void **out) {
  if (bytes > ctx->scratch_bytes[slot]) {
    if (ctx->scratch[slot]) ZC_CUDA(ctx, cudaFree(ctx->scratch[slot]));
    ctx->scratch[slot] = nullptr;
    ctx->scratch_bytes[slot] = 0;
    size_t want = bytes + (bytes >> 3) + 256;
    ZC_CUDA(ctx, cudaMalloc(&ctx->scratch[slot], want));
    ctx->scratch_bytes[slot] = want;
  }
  *out = ctx->scratch[slot];
  return ZC_OK;
}

// implemented in zc_peer.cu
int32_t zc_peer_exchange_fold(zc_ctx *ctx, const uint64_t *partial, uint64_t *out);
int32_t zc_peer_check_error(zc_ctx *ctx);
// implemented in zc_kernels.cu (extern "C", not in the public header): enqueue a canonical-input check (kind 1 field, 2 scalar, 3 point) / read the verdict
int32_t zc_validate_dev(zc_ctx *ctx, int32_t kind, const uint64_t *a, size_t n, size_t base);
int32_t zc_validate_finish(zc_ctx *ctx, int32_t elems_per_unit);
// implemented in zc_msm.cu
int32_t zc_msm_run(zc_ctx *ctx, const zc_msm_generators *gens, const uint64_t *points, const uint64_t *scalars, size_t n,
                   int32_t window_bits, int32_t rank, int32_t nranks, bool exchange, uint64_t *out_point_dev);
