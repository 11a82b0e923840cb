// NCCL communicator handle of the multi-GPU entry points (include/fvgp_b200.h, "multi-GPU").
//
// The library does not link NCCL: the process that calls it (one rank per GPU under torchrun) has already
// loaded the NCCL that torch ships, and two different NCCL builds in one process must be avoided.
// fvgp_nccl_attach() therefore dlopen()s the shared object the caller names (or the SONAME, which resolves to
// the copy that is already mapped) and binds the handful of entry points used here.  The handle wraps either a
// communicator created by fvgp_comm_create (ncclCommInitRank on a unique id the host side distributes) or an
// existing ncclComm_t the caller owns (fvgp_comm_adopt).
#pragma once
#include "common.cuh"
#include <nccl.h>

namespace fvgp {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi& nccl_api();  // comm.cu

struct Comm {
  ncclComm_t comm;
  int rank, world;
  bool owned;
};

#define FVGP_NCCL_OK(expr)                                                                                   \
  do {                                                                                                       \
    ncclResult_t _r = (expr);                                                                                \
    if (_r != ncclSuccess) {                                                                                 \
      fprintf(stderr, "[fvgp_b200] NCCL error %d at %s:%d: %s\n", (int)_r, __FILE__, __LINE__,               \
              fvgp::nccl_api().GetErrorString ? fvgp::nccl_api().GetErrorString(_r) : "?");                  \
      return FVGP_ERR_CUDA;                                                                                  \
    }                                                                                                        \
  } while (0)

// In-place all-gather of unequal parts: rank r owns bytes [off[r], off[r+1]) of d_buf.  One grouped launch.
inline int comm_allgatherv_bytes(const Comm* c, void* d_buf, const int64_t* h_off, cudaStream_t st) {
  NcclApi& api = nccl_api();
  if (c->world == 1) return 0;
  FVGP_NCCL_OK(api.GroupStart());
  for (int r = 0; r < c->world; ++r) {
    const int64_t len = h_off[r + 1] - h_off[r];
    if (len <= 0) continue;
    char* at = (char*)d_buf + h_off[r];
    ncclResult_t res = (len % 8 == 0 && ((uintptr_t)at) % 8 == 0)
                           ? api.Broadcast(at, at, (size_t)(len / 8), ncclDouble, r, c->comm, st)
                           : api.Broadcast(at, at, (size_t)len, ncclInt8, r, c->comm, st);
    if (res != ncclSuccess) {
      api.GroupEnd();
      FVGP_NCCL_OK(res);
    }
  }
  FVGP_NCCL_OK(api.GroupEnd());
  return 0;
}

inline int comm_allreduce_sum(const Comm* c, double* d_buf, size_t count, cudaStream_t st) {
  if (c->world == 1) return 0;
  FVGP_NCCL_OK(nccl_api().AllReduce(d_buf, d_buf, count, ncclDouble, ncclSum, c->comm, st));
  return 0;
}

}  // namespace fvgp
