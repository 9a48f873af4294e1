// The one collective of the path: the data-parallel gradient all-reduce -- plus the optional
// BatchNorm-statistics all-reduce of sync-BN -- behind the C ABI.  NCCL is bound at run time with dlopen("libnccl.so.2"):
// inside a process that already imported torch this resolves to the very library torch loaded (same SONAME), so the
// communicator shares NVLink / NVSwitch transport state with nothing else and the .so itself has no link-time NCCL
// dependency (it still loads, and everything but hgb_comm_* works, on a machine without NCCL).
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"
#include "comm.cuh"

namespace hgb {

namespace {
struct NcclApi {
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
  bool ok = false;
};

NcclApi* nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.ok ? &api : nullptr;
  tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return nullptr;
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
  api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
  api.GetVersion = (decltype(api.GetVersion))dlsym(h, "ncclGetVersion");
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString;
  return api.ok ? &api : nullptr;
}

#define HGB_NCCL(expr)                                                                                   \
  do {                                                                                                   \
    ncclResult_t _r = (expr);                                                                            \
    if (_r != ncclSuccess) {                                                                             \
      set_error("%s failed: %s (%s:%d)", #expr, nccl()->GetErrorString(_r), __FILE__, __LINE__);         \
      return HGB_ERR_CUDA;                                                                               \
    }                                                                                                    \
  } while (0)
}  // namespace

int comm_allreduce_sum_f32(hgb_comm* c, float* buf, int64_t count, cudaStream_t st) {
  HGB_CHECK_ARG(c && c->handle, "all-reduce: no communicator");
  if (count == 0) return HGB_OK;
  HGB_NCCL(nccl()->AllReduce(buf, buf, (size_t)count, ncclFloat32, ncclSum, (ncclComm_t)c->handle, st));
  return HGB_OK;
}

}  // namespace hgb

using namespace hgb;

extern "C" int hgb_comm_unique_id(void* id_out, int bytes) {
  HGB_CHECK_ARG(id_out && bytes >= NCCL_UNIQUE_ID_BYTES, "hgb_comm_unique_id: need a buffer of %d bytes", NCCL_UNIQUE_ID_BYTES);
  if (!nccl()) { set_error("hgb_comm_unique_id: libnccl.so.2 cannot be loaded"); return HGB_ERR_STATE; }
  ncclUniqueId id;
  HGB_NCCL(nccl()->GetUniqueId(&id));
  memcpy(id_out, &id, NCCL_UNIQUE_ID_BYTES);
  return HGB_OK;
}

extern "C" int hgb_comm_init(int nranks, int rank, const void* unique_id, hgb_comm** out) {
  HGB_CHECK_ARG(out && unique_id && nranks >= 1 && rank >= 0 && rank < nranks, "hgb_comm_init: bad arguments");
  if (!nccl()) { set_error("hgb_comm_init: libnccl.so.2 cannot be loaded"); return HGB_ERR_STATE; }
  ncclUniqueId id;
  memcpy(&id, unique_id, NCCL_UNIQUE_ID_BYTES);
  ncclComm_t comm = nullptr;
  HGB_NCCL(nccl()->CommInitRank(&comm, nranks, id, rank));
  hgb_comm* c = new hgb_comm();
  c->handle = comm; c->nranks = nranks; c->rank = rank;
  *out = c;
  return HGB_OK;
}

extern "C" int hgb_comm_destroy(hgb_comm* c) {
  if (c) {
    if (c->handle && nccl()) nccl()->CommDestroy((ncclComm_t)c->handle);
    delete c;
  }
  return HGB_OK;
}

extern "C" int hgb_comm_info(const hgb_comm* c, int* nranks, int* rank, int* nccl_version) {
  HGB_CHECK_ARG(c, "hgb_comm_info: null communicator");
  if (nranks) *nranks = c->nranks;
  if (rank) *rank = c->rank;
  if (nccl_version) {
    int v = 0;
    if (nccl() && nccl()->GetVersion) nccl()->GetVersion(&v);
    *nccl_version = v;
  }
  return HGB_OK;
}

extern "C" int hgb_comm_allreduce_f32(hgb_comm* c, float* buf, int64_t count, void* stream) {
  HGB_CHECK_ARG(buf || count == 0, "hgb_comm_allreduce_f32: null buffer");
  return comm_allreduce_sum_f32(c, buf, count, (cudaStream_t)stream);
}
