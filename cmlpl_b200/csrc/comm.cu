// Collective entry points of the C ABI (comm_init / allgather_labels / allreduce_confusion) for hosts that
// do not bring torch.distributed: the two exchanges at the end of a band-sharded scene inference
// (tools/hyper_tools.py:416-437 over row bands, :208-223 confusion counts), on NCCL over NVLink.
// NCCL is resolved at run time (dlopen of libnccl.so.2, the copy already loaded by the process if there is one), so
// libcmlpl_sm100.so has no link-time dependency on it; the Python layer keeps using torch.distributed.
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace cmlpl {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) {
      api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.handle, "ncclGetUniqueId"));
      api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.handle, "ncclCommInitRank"));
      api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.handle, "ncclAllGather"));
      api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(api.handle, "ncclAllReduce"));
      api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.handle, "ncclCommDestroy"));
      api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.handle, "ncclGetErrorString"));
    }
  }
  const bool ok = api.handle && api.GetUniqueId && api.CommInitRank && api.AllGather && api.AllReduce && api.CommDestroy;
  return ok ? &api : nullptr;
}

}  // namespace cmlpl

struct cmlpl_comm {
  ncclComm_t comm;
  int rank, world;
};

using namespace cmlpl;

#define CMLPL_NCCL(api, call, what)                                                                            \
  do {                                                                                                         \
    ncclResult_t r__ = (call);                                                                                 \
    if (r__ != ncclSuccess) {                                                                                  \
      set_error("%s failed: %s", what, (api)->GetErrorString ? (api)->GetErrorString(r__) : "NCCL error");     \
      return CMLPL_ERR_CUDA;                                                                                   \
    }                                                                                                          \
  } while (0)

extern "C" int cmlpl_comm_unique_id(void* id128) {
  CMLPL_CHECK_ARG(id128, "comm_unique_id: null pointer");
  NcclApi* api = nccl_api();
  CMLPL_CHECK_ARG(api, "comm_unique_id: libnccl.so.2 could not be loaded");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  CMLPL_NCCL(api, api->GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(id128, &id, sizeof(id));
  return CMLPL_OK;
}

extern "C" int cmlpl_comm_init(int rank, int world, const void* id128, cmlpl_comm** out) {
  CMLPL_CHECK_ARG(id128 && out && world >= 1 && rank >= 0 && rank < world, "comm_init: bad arguments");
  NcclApi* api = nccl_api();
  CMLPL_CHECK_ARG(api, "comm_init: libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  cmlpl_comm* c = new cmlpl_comm{nullptr, rank, world};
  ncclResult_t r = api->CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) {
    set_error("ncclCommInitRank failed: %s", api->GetErrorString ? api->GetErrorString(r) : "NCCL error");
    delete c;
    return CMLPL_ERR_CUDA;
  }
  *out = c;
  return CMLPL_OK;
}

extern "C" int cmlpl_comm_allgather_labels(cmlpl_comm* c, const uint8_t* local, int64_t per_rank, uint8_t* out,
                                           cmlpl_stream_t stream) {
  CMLPL_CHECK_ARG(c && local && out && per_rank >= 0, "comm_allgather_labels: bad arguments");
  NcclApi* api = nccl_api();
  CMLPL_CHECK_ARG(api, "comm_allgather_labels: NCCL not loaded");
  if (per_rank == 0) return CMLPL_OK;
  CMLPL_NCCL(api, api->AllGather(local, out, size_t(per_rank), ncclUint8, c->comm, static_cast<cudaStream_t>(stream)),
             "ncclAllGather");
  return CMLPL_OK;
}

extern "C" int cmlpl_comm_allreduce_confusion(cmlpl_comm* c, int64_t* cm, int count, cmlpl_stream_t stream) {
  CMLPL_CHECK_ARG(c && cm && count > 0, "comm_allreduce_confusion: bad arguments");
  NcclApi* api = nccl_api();
  CMLPL_CHECK_ARG(api, "comm_allreduce_confusion: NCCL not loaded");
  CMLPL_NCCL(api, api->AllReduce(cm, cm, size_t(count), ncclInt64, ncclSum, c->comm, static_cast<cudaStream_t>(stream)),
             "ncclAllReduce");
  return CMLPL_OK;
}

extern "C" int cmlpl_comm_destroy(cmlpl_comm* c) {
  if (!c) return CMLPL_OK;
  NcclApi* api = nccl_api();
  if (api && c->comm) api->CommDestroy(c->comm);
  delete c;
  return CMLPL_OK;
}
