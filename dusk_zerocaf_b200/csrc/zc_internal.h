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
  int32_t c, rank, nranks, pad;
  const void *partial, *ws;
};

// NVLink peer-memory exchange (zc_peer.cu): one mailbox per rank, mapped into every peer with CUDA IPC
#define ZC_MAX_PEERS 16
struct zc_mailbox {
  uint64_t flag[ZC_MAX_PEERS];        // flag[r] = sequence number of the last partial rank r delivered here
  uint64_t counter;                   // this rank's own exchange count
  uint64_t pad[15];
  uint64_t slot[ZC_MAX_PEERS][20];    // slot[r] = rank r's partial point, [u64;20]
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
  // host-pointer entry points: copy-in / copy-out streams and per-chunk events (created on first use)
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  cudaEvent_t pipe_ev[2 * ZC_PIPE_MAX_CHUNKS] = {};
  // MSM operands prepared once for fixed generators (zc_msm_prepare_points_dev)
  const void *prep_points = nullptr;
  size_t prep_n = 0;
  bool prep_valid = false;
  // fixed-base MSM tables (zc_msm_prepare_fixed_base_dev): 2^(c w) P_i for this rank's windows, affine cached form
  void *fb_table = nullptr;
  void *fb_corr = nullptr;          // [u64;20]: minus the constant a spread short window adds (zc_msm.cu), or null
  size_t fb_table_bytes = 0;
  const void *fb_points = nullptr;
  size_t fb_n = 0;
  int32_t fb_c = 0, fb_rank = 0, fb_nranks = 0;
  void *basepoint_table = nullptr;  // fixed-base table (zc_fixed.cu), built on first use
  void *msm_graph_exec = nullptr;   // cudaGraphExec_t of the last MSM configuration
  zc_msm_key msm_key = {};
  uint64_t msm_graph_launches = 0;
  // MSM: side stream for the window-scaling chain + events (created on first use)
  cudaStream_t side_stream = nullptr, side_extra[3] = {nullptr, nullptr, nullptr}, chain_stream = nullptr, sort_stream = nullptr;
  cudaStream_t side_hi[3] = {nullptr, nullptr, nullptr};   // high-priority twins of the first three side streams (sharded MSM)
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
// implemented in zc_msm.cu
int32_t zc_msm_run(zc_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, int32_t window_bits,
                   int32_t rank, int32_t nranks, bool exchange, uint64_t *out_point_dev);
