// K8: the one collective of the job -- every rank's per-(run, agent) metric block and per-run revenue to every rank
// (reference src/main.py:186-222 keeps one row per run, agent and iteration, so the ranks' blocks are gathered, not
// reduced).  NCCL is bound at run time (dlopen of libnccl.so.2: the copy torch has already loaded when the caller is the
// Python host, the system copy otherwise), so libagym.so has no link-time dependency on it and a single-GPU user never
// touches it.  One communicator per handle; the unique id travels through whatever the host already has (torch.distributed
// broadcast_object_list in auction_gym_b200/engine.py, MPI_Bcast in a C host).
#include <dlfcn.h>

#include <cstring>

#include "agym_common.cuh"

namespace {

struct NcclId {
  char internal[128];  // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES 128)
};
using ncclComm_t = void*;
constexpr int kNcclFloat64 = 8;  // ncclDouble / ncclFloat64 (nccl.h ncclDataType_t)

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclId, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string err;
};

NcclApi& nccl() {
  static NcclApi api;
  if (api.lib || !api.err.empty()) return api;
  api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!api.lib) { api.err = std::string("libnccl.so.2 not found: ") + dlerror(); return api; }
  auto sym = [&](const char* n) { void* p = dlsym(api.lib, n); if (!p) api.err = std::string("NCCL symbol missing: ") + n; return p; };
  api.GetUniqueId = reinterpret_cast<int (*)(NcclId*)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<int (*)(ncclComm_t*, int, NcclId, int)>(sym("ncclCommInitRank"));
  api.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t)>(sym("ncclAllGather"));
  api.CommDestroy = reinterpret_cast<int (*)(ncclComm_t)>(sym("ncclCommDestroy"));
  api.GroupStart = reinterpret_cast<int (*)()>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<int (*)()>(sym("ncclGroupEnd"));
  api.GetErrorString = reinterpret_cast<const char* (*)(int)>(sym("ncclGetErrorString"));
  return api;
}

int nccl_fail(agym_handle* h, int rc, const char* what) {
  NcclApi& n = nccl();
  return agym::set_error(h, AGYM_ERR_CUDA, std::string(what) + ": NCCL error " + std::to_string(rc) + " (" + (n.GetErrorString ? n.GetErrorString(rc) : "?") + ")");
}

}  // namespace

namespace agym {
void destroy_comm(agym_handle* h) {
  if (h->nccl_comm && nccl().CommDestroy) nccl().CommDestroy(h->nccl_comm);
  h->nccl_comm = nullptr;
}
}  // namespace agym

extern "C" {

int agym_nccl_unique_id(char* out128) {
  if (!out128) return AGYM_ERR_INVALID;
  NcclApi& n = nccl();
  if (!n.err.empty()) return agym::set_error(nullptr, AGYM_ERR_STATE, n.err);
  NcclId id;
  const int rc = n.GetUniqueId(&id);
  if (rc) return nccl_fail(nullptr, rc, "agym_nccl_unique_id");
  std::memcpy(out128, id.internal, sizeof(id.internal));
  return AGYM_OK;
}

int agym_comm_init(agym_handle* h, const char* id128, int32_t rank, int32_t world) {
  if (!h || !id128 || world < 1 || rank < 0 || rank >= world) return agym::set_error(h, AGYM_ERR_INVALID, "agym_comm_init: bad argument");
  NcclApi& n = nccl();
  if (!n.err.empty()) return agym::set_error(h, AGYM_ERR_STATE, n.err);
  agym::destroy_comm(h);
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(h->device);
  NcclId id;
  std::memcpy(id.internal, id128, sizeof(id.internal));
  ncclComm_t comm = nullptr;
  const int rc = n.CommInitRank(&comm, world, id, rank);
  if (prev >= 0 && prev != h->device) cudaSetDevice(prev);
  if (rc) return nccl_fail(h, rc, "agym_comm_init");
  h->nccl_comm = comm;
  h->nccl_rank = rank;
  h->nccl_world = world;
  return AGYM_OK;
}

int agym_gather_metrics_nccl(agym_handle* h, double* recv_acc, double* recv_revenue, void* stream) {
  if (!h || !recv_acc || !recv_revenue) return agym::set_error(h, AGYM_ERR_INVALID, "agym_gather_metrics_nccl: null argument");
  if (!h->nccl_comm) return agym::set_error(h, AGYM_ERR_STATE, "agym_gather_metrics_nccl: call agym_comm_init first");
  if (!h->acc || !h->revenue) return agym::set_error(h, AGYM_ERR_STATE, "agym_gather_metrics_nccl: metrics not bound");
  NcclApi& n = nccl();
  const agym_shape& s = h->shape;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = n.GroupStart();
  if (!rc) rc = n.AllGather(h->acc, recv_acc, (size_t)s.R * s.A * AGYM_NUM_METRICS, kNcclFloat64, h->nccl_comm, st);
  if (!rc) rc = n.AllGather(h->revenue, recv_revenue, (size_t)s.R, kNcclFloat64, h->nccl_comm, st);
  const int rc2 = n.GroupEnd();
  if (rc || rc2) return nccl_fail(h, rc ? rc : rc2, "agym_gather_metrics_nccl");
  return AGYM_OK;
}

int agym_gather_block_nccl(agym_handle* h, const double* send, double* recv, int64_t count, void* stream) {
  if (!h || !send || !recv || count < 0) return agym::set_error(h, AGYM_ERR_INVALID, "agym_gather_block_nccl: bad argument");
  if (!h->nccl_comm) return agym::set_error(h, AGYM_ERR_STATE, "agym_gather_block_nccl: call agym_comm_init first");
  if (count == 0) return AGYM_OK;
  const int rc = nccl().AllGather(send, recv, (size_t)count, kNcclFloat64, h->nccl_comm, (cudaStream_t)stream);
  if (rc) return nccl_fail(h, rc, "agym_gather_block_nccl");
  return AGYM_OK;
}

}  // extern "C"
