// Data-parallel exchange step (the reference, leosampaio/sketchformer, is single-device and has no
// collective): gradient buckets of the flat arena all-reduced with NCCL over NVLink / NVSwitch.
//
// NCCL is bound at RUN time (dlopen): libskf.so carries no link-time dependency on it, single-GPU users never load
// it, and inside a torch process the library torch already loaded (its bundled libnccl.so.2) is the one that is used,
// so there is exactly one NCCL in the process.  Only the stable core API is touched: ncclGetUniqueId,
// ncclCommInitRank, ncclAllReduce, ncclCommDestroy, ncclGetErrorString.
#include <dlfcn.h>

#include <mutex>

#include "skf_common.cuh"

namespace skf {

// ---- the few NCCL declarations needed (nccl.h, stable since 2.0) -----------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;                 // ncclSuccess = 0
enum { NCCL_FLOAT32 = 7, NCCL_SUM = 0 };  // ncclDataType_t::ncclFloat32, ncclRedOp_t::ncclSum

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  char why[256] = {0};
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // 1. the copy already mapped into this process (torch's), 2. the regular search path
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (int pass = 0; pass < 2 && !api.handle; pass++)
      for (const char* n : names) {
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL | (pass == 0 ? RTLD_NOLOAD : 0));
        if (api.handle) break;
      }
    if (!api.handle) {
      snprintf(api.why, sizeof(api.why), "libnccl.so.2 not found (%s)", dlerror());
      return;
    }
    bool ok = true;
    auto sym = [&](const char* s) {
      void* p = dlsym(api.handle, s);
      if (!p) { ok = false; snprintf(api.why, sizeof(api.why), "symbol %s missing in libnccl", s); }
      return p;
    };
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    if (!ok) api.handle = nullptr;
  });
  return &api;
}

#define SKF_NCCL(api, expr)                                                                              \
  do {                                                                                                   \
    ncclResult_t _r = (expr);                                                                            \
    if (_r != 0)                                                                                         \
      return skf::fail(SKF_ERR_NCCL, "%s:%d %s: %s", __FILE__, __LINE__, #expr, (api)->GetErrorString(_r)); \
  } while (0)

int allreduce_sum_f32(float* ptr, int64_t count, void* comm, cudaStream_t st) {
  NcclApi* api = nccl_api();
  SKF_REQUIRE(api->handle != nullptr, SKF_ERR_NCCL, "NCCL is not available: %s", api->why);
  SKF_REQUIRE(ptr != nullptr && comm != nullptr && count > 0, SKF_ERR_BAD_ARG, "allreduce_bucket: bad argument");
  SKF_NCCL(api, api->AllReduce(ptr, ptr, (size_t)count, NCCL_FLOAT32, NCCL_SUM, reinterpret_cast<ncclComm_t>(comm), st));
  return SKF_OK;
}

}  // namespace skf

using namespace skf;

extern "C" {

int skf_nccl_version(void) {
  NcclApi* api = nccl_api();
  if (!api->handle) return fail(SKF_ERR_NCCL, "NCCL is not available: %s", api->why);
  int v = 0;
  SKF_NCCL(api, api->GetVersion(&v));
  return v;
}

int skf_comm_unique_id(void* id128) {
  NcclApi* api = nccl_api();
  SKF_REQUIRE(api->handle != nullptr, SKF_ERR_NCCL, "NCCL is not available: %s", api->why);
  SKF_REQUIRE(id128 != nullptr, SKF_ERR_BAD_ARG, "comm_unique_id: NULL buffer");
  ncclUniqueId id;
  SKF_NCCL(api, api->GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return SKF_OK;
}

int skf_comm_create(const void* id128, int world_size, int rank, void** comm) {
  NcclApi* api = nccl_api();
  SKF_REQUIRE(api->handle != nullptr, SKF_ERR_NCCL, "NCCL is not available: %s", api->why);
  SKF_REQUIRE(id128 && comm && world_size >= 1 && rank >= 0 && rank < world_size, SKF_ERR_BAD_ARG,
              "comm_create: bad argument (world %d, rank %d)", world_size, rank);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c = nullptr;
  SKF_NCCL(api, api->CommInitRank(&c, world_size, id, rank));
  *comm = c;
  return SKF_OK;
}

int skf_comm_destroy(void* comm) {
  if (!comm) return SKF_OK;
  NcclApi* api = nccl_api();
  SKF_REQUIRE(api->handle != nullptr, SKF_ERR_NCCL, "NCCL is not available: %s", api->why);
  SKF_NCCL(api, api->CommDestroy(reinterpret_cast<ncclComm_t>(comm)));
  return SKF_OK;
}

int skf_allreduce_bucket(float* ptr, int64_t count, void* comm, void* stream) {
  return allreduce_sum_f32(ptr, count, comm, as_stream(stream));
}

}  // extern "C"
