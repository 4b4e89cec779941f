// zc_nccl.cu -- the one exchange step of the bucket-window-sharded MSM: an all-gather of the 160-byte partial points.
//
// libnccl is resolved at run time (dlopen), so libzerocaf_b200.so loads on hosts without NCCL and, inside a torch
// process, binds to the NCCL that torch already loaded.  Elliptic-curve addition is not an ncclRedOp_t, so the
// "all-reduce of partial sums" is all-gather + a fixed-order fold with the reference's Add on every rank (zc_msm.cu).
#include <dlfcn.h>

#include "zc_internal.h"

namespace {

typedef struct { char internal[128]; } zc_ncclUniqueId;
typedef int (*fn_GetUniqueId)(zc_ncclUniqueId*);
typedef int (*fn_CommInitRank)(void**, int, zc_ncclUniqueId, int);
typedef int (*fn_CommDestroy)(void*);
typedef int (*fn_AllGather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*fn_GetErrorString)(int);

struct NcclApi {
  void* handle = nullptr;
  fn_GetUniqueId GetUniqueId = nullptr;
  fn_CommInitRank CommInitRank = nullptr;
  fn_CommDestroy CommDestroy = nullptr;
  fn_AllGather AllGather = nullptr;
  fn_GetErrorString GetErrorString = nullptr;
  bool ok = false;
};

NcclApi& api() {
  static NcclApi a;
  if (a.handle) return a;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    a.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (a.handle) break;
  }
  if (!a.handle) return a;
  a.GetUniqueId = (fn_GetUniqueId)dlsym(a.handle, "ncclGetUniqueId");
  a.CommInitRank = (fn_CommInitRank)dlsym(a.handle, "ncclCommInitRank");
  a.CommDestroy = (fn_CommDestroy)dlsym(a.handle, "ncclCommDestroy");
  a.AllGather = (fn_AllGather)dlsym(a.handle, "ncclAllGather");
  a.GetErrorString = (fn_GetErrorString)dlsym(a.handle, "ncclGetErrorString");
  a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather;
  return a;
}

constexpr int32_t ZC_NCCL_FAIL = -100000;   // NCCL result r is reported as ZC_NCCL_FAIL - r

}  // namespace

int32_t zc_nccl_allgather(zc_ctx* ctx, const void* send, void* recv, size_t bytes) {
  NcclApi& a = api();
  if (!a.ok) return zc_fail(ctx, ZC_ERR_STATE, "libnccl not available");
  int r = a.AllGather(send, recv, bytes, /*ncclUint8*/ 1, ctx->nccl_comm, ctx->stream);
  if (r != 0) {
    snprintf(ctx->err, sizeof(ctx->err), "ncclAllGather -> %s", a.GetErrorString ? a.GetErrorString(r) : "error");
    return ZC_NCCL_FAIL - r;
  }
  return ZC_OK;
}

extern "C" {

int32_t zc_nccl_unique_id(uint8_t id_out[128]) {
  if (!id_out) return ZC_ERR_NULL;
  NcclApi& a = api();
  if (!a.ok) return ZC_ERR_STATE;
  zc_ncclUniqueId id;
  int r = a.GetUniqueId(&id);
  if (r != 0) return ZC_NCCL_FAIL - r;
  memcpy(id_out, id.internal, 128);
  return ZC_OK;
}

int32_t zc_nccl_comm_init(const uint8_t id_in[128], int32_t rank, int32_t nranks, void** comm_out) {
  if (!id_in || !comm_out) return ZC_ERR_NULL;
  NcclApi& a = api();
  if (!a.ok) return ZC_ERR_STATE;
  zc_ncclUniqueId id;
  memcpy(id.internal, id_in, 128);
  void* comm = nullptr;
  int r = a.CommInitRank(&comm, nranks, id, rank);
  if (r != 0) return ZC_NCCL_FAIL - r;
  *comm_out = comm;
  return ZC_OK;
}

int32_t zc_nccl_comm_destroy(void* comm) {
  NcclApi& a = api();
  if (!a.ok) return ZC_ERR_STATE;
  int r = a.CommDestroy(comm);
  return r == 0 ? ZC_OK : ZC_NCCL_FAIL - r;
}

int32_t zc_ctx_set_nccl(zc_ctx* ctx, void* nccl_comm, int32_t rank, int32_t nranks) {
  if (!ctx) return ZC_ERR_NULL;
  if (nranks < 1 || rank < 0 || rank >= nranks) return zc_fail(ctx, ZC_ERR_SIZE, "bad rank / nranks");
  ZC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ctx->gather_buf) { ZC_CUDA(ctx, cudaFree(ctx->gather_buf)); ctx->gather_buf = nullptr; }
  ZC_CUDA(ctx, cudaMalloc(&ctx->gather_buf, (size_t)(nranks + 1) * 160));
  ctx->nccl_comm = nccl_comm;
  ctx->rank = rank;
  ctx->nranks = nranks;
  return ZC_OK;
}

}  // extern "C"
