// Communicator management of the multi-GPU C ABI (see comm.cuh, include/fvgp_b200.h).
#include "../../include/fvgp_b200.h"
#include "comm.cuh"
#include <dlfcn.h>
#include <string.h>

namespace fvgp {

NcclApi& nccl_api() {
  static NcclApi api;
  return api;
}

template <typename F>
static bool bind(void* h, const char* name, F& fn) {
  fn = reinterpret_cast<F>(dlsym(h, name));
  if (fn == nullptr) fprintf(stderr, "[fvgp_b200] NCCL symbol %s not found\n", name);
  return fn != nullptr;
}

}  // namespace fvgp

using namespace fvgp;

extern "C" {

int fvgp_nccl_attach(const char* libnccl_path) {
  NcclApi& api = nccl_api();
  if (api.handle != nullptr) return 0;
  const char* path = (libnccl_path != nullptr && libnccl_path[0] != 0) ? libnccl_path : "libnccl.so.2";
  void* h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) {
    fprintf(stderr, "[fvgp_b200] cannot load NCCL (%s): %s\n", path, dlerror());
    return FVGP_ERR_ARG;
  }
  bool ok = bind(h, "ncclGetUniqueId", api.GetUniqueId) & bind(h, "ncclCommInitRank", api.CommInitRank) &
            bind(h, "ncclCommDestroy", api.CommDestroy) & bind(h, "ncclAllReduce", api.AllReduce) &
            bind(h, "ncclBroadcast", api.Broadcast) & bind(h, "ncclGroupStart", api.GroupStart) &
            bind(h, "ncclGroupEnd", api.GroupEnd) & bind(h, "ncclGetErrorString", api.GetErrorString);
  if (!ok) return FVGP_ERR_ARG;
  api.handle = h;
  return 0;
}

int fvgp_comm_unique_id(void* h_id128) {
  FVGP_REQUIRE(nccl_api().handle != nullptr && h_id128 != nullptr);
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  FVGP_NCCL_OK(nccl_api().GetUniqueId(&id));
  memcpy(h_id128, &id, sizeof(id));
  return 0;
}

int fvgp_comm_create(const void* h_id128, int rank, int world, void** comm_out) {
  FVGP_REQUIRE(nccl_api().handle != nullptr && h_id128 != nullptr && comm_out != nullptr && world >= 1 && rank >= 0 &&
               rank < world);
  ncclUniqueId id;
  memcpy(&id, h_id128, sizeof(id));
  Comm* c = new Comm{nullptr, rank, world, true};
  ncclResult_t r = nccl_api().CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) {
    delete c;
    FVGP_NCCL_OK(r);
  }
  *comm_out = c;
  return 0;
}

int fvgp_comm_adopt(void* nccl_comm, int rank, int world, void** comm_out) {
  FVGP_REQUIRE(nccl_api().handle != nullptr && nccl_comm != nullptr && comm_out != nullptr && world >= 1 && rank >= 0 &&
               rank < world);
  *comm_out = new Comm{(ncclComm_t)nccl_comm, rank, world, false};
  return 0;
}

int fvgp_comm_destroy(void* comm) {
  if (comm == nullptr) return 0;
  Comm* c = (Comm*)comm;
  if (c->owned && c->comm != nullptr && nccl_api().CommDestroy) nccl_api().CommDestroy(c->comm);
  delete c;
  return 0;
}

int fvgp_comm_allgatherv(void* comm, void* d_buf, const int64_t* h_offsets_bytes, void* stream) {
  FVGP_REQUIRE(comm != nullptr && d_buf != nullptr && h_offsets_bytes != nullptr);
  return comm_allgatherv_bytes((const Comm*)comm, d_buf, h_offsets_bytes, (cudaStream_t)stream);
}

int fvgp_comm_allreduce_sum(void* comm, double* d_buf, int64_t count, void* stream) {
  FVGP_REQUIRE(comm != nullptr && d_buf != nullptr && count >= 0);
  if (count == 0) return 0;
  return comm_allreduce_sum((const Comm*)comm, d_buf, (size_t)count, (cudaStream_t)stream);
}

}  // extern "C"
