// Collectives for a host that is not Python (include/chromegcn.h, "collectives for a non-Python host").
//
// The three exchanges the multi-GPU paths need -- flat-gradient sum (chromosome-sharded pass), BatchNorm column sums
// (row-partitioned graph) and the panel all-gather of the copy-based exchange -- as thin calls over NCCL.  NCCL is
// resolved with dlopen at first use: libchromegcn.so has no link-time NCCL dependency, a process that already mapped a
// libnccl.so.2 (PyTorch's bundled copy, for one) gets that copy, and a box without NCCL gets an error message instead of
// a loader failure.  Only the handful of NCCL 2 symbols below are used; their signatures and enum values have been
// stable across NCCL 2.x, so they are declared here rather than taken from nccl.h.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace cgcn {
namespace {

struct NcclId {
  char bytes[CGCN_COMM_ID_BYTES];                  // ncclUniqueId: NCCL_UNIQUE_ID_BYTES == 128
};
typedef void* NcclComm;                            // ncclComm_t
constexpr int kNcclSuccess = 0;
constexpr int kNcclInt8 = 0, kNcclFloat32 = 7;     // ncclDataType_t
constexpr int kNcclSum = 0;                        // ncclRedOp_t

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  char why[256] = {0};
};

NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {getenv("CGCN_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* name : names) {
      if (name == nullptr || *name == 0) continue;
      api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle != nullptr) break;
      snprintf(api.why, sizeof(api.why), "%s", dlerror());
    }
    if (api.handle == nullptr) return;
    bool ok = true;
    auto sym = [&](const char* s) {
      void* p = dlsym(api.handle, s);
      if (p == nullptr) {
        ok = false;
        snprintf(api.why, sizeof(api.why), "symbol %s not found in NCCL", s);
      }
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    if (!ok) {
      dlclose(api.handle);
      api.handle = nullptr;
    }
  });
  return &api;
}

int need_nccl(NcclApi** out) {
  NcclApi* api = nccl_api();
  if (api->handle == nullptr) {
    set_error("cgcn_comm: NCCL is not available (%s); set CGCN_NCCL_LIB to a libnccl.so.2", api->why[0] ? api->why : "not found");
    return CGCN_ERR_INVALID;
  }
  *out = api;
  return CGCN_OK;
}

int nccl_status(NcclApi* api, int rc, const char* what) {
  if (rc == kNcclSuccess) return CGCN_OK;
  set_error("%s: NCCL error %d (%s)", what, rc, api->GetErrorString ? api->GetErrorString(rc) : "?");
  return CGCN_ERR_CUDA;
}

}  // namespace
}  // namespace cgcn

struct cgcn_comm {
  cgcn::NcclComm nccl;
  int32_t world, rank, device;
};

extern "C" int cgcn_comm_unique_id(unsigned char id_host[CGCN_COMM_ID_BYTES]) {
  using namespace cgcn;
  CGCN_REQUIRE(id_host != nullptr, "cgcn_comm_unique_id: null id");
  NcclApi* api = nullptr;
  CGCN_TRY(need_nccl(&api));
  NcclId id;
  CGCN_TRY(nccl_status(api, api->GetUniqueId(&id), "ncclGetUniqueId"));
  memcpy(id_host, id.bytes, CGCN_COMM_ID_BYTES);
  return CGCN_OK;
}

extern "C" int cgcn_comm_init(cgcn_comm_t* comm_host, const unsigned char id_host[CGCN_COMM_ID_BYTES], int32_t world, int32_t rank) {
  using namespace cgcn;
  CGCN_REQUIRE(comm_host != nullptr && id_host != nullptr, "cgcn_comm_init: null argument");
  CGCN_REQUIRE(world >= 1 && rank >= 0 && rank < world, "cgcn_comm_init: rank %d of world %d", rank, world);
  *comm_host = nullptr;
  NcclApi* api = nullptr;
  CGCN_TRY(need_nccl(&api));
  int dev = 0;
  CGCN_CUDA(cudaGetDevice(&dev));
  NcclId id;
  memcpy(id.bytes, id_host, CGCN_COMM_ID_BYTES);
  NcclComm c = nullptr;
  CGCN_TRY(nccl_status(api, api->CommInitRank(&c, world, id, rank), "ncclCommInitRank"));
  cgcn_comm* out = new cgcn_comm{c, world, rank, dev};
  *comm_host = out;
  return CGCN_OK;
}

extern "C" int cgcn_comm_destroy(cgcn_comm_t comm) {
  using namespace cgcn;
  if (comm == nullptr) return CGCN_OK;
  NcclApi* api = nullptr;
  CGCN_TRY(need_nccl(&api));
  const int rc = api->CommDestroy(comm->nccl);
  delete comm;
  return nccl_status(api, rc, "ncclCommDestroy");
}

extern "C" int cgcn_comm_allreduce_sum(cgcn_comm_t comm, float* buf, size_t count, cgcn_stream_t stream) {
  using namespace cgcn;
  CGCN_REQUIRE(comm != nullptr && (buf != nullptr || count == 0), "cgcn_comm_allreduce_sum: null argument");
  if (count == 0) return CGCN_OK;
  NcclApi* api = nullptr;
  CGCN_TRY(need_nccl(&api));
  return nccl_status(api, api->AllReduce(buf, buf, count, kNcclFloat32, kNcclSum, comm->nccl, static_cast<cudaStream_t>(stream)),
                     "ncclAllReduce");
}

extern "C" int cgcn_comm_allgather(cgcn_comm_t comm, const void* send, void* recv, size_t bytes_per_rank, cgcn_stream_t stream) {
  using namespace cgcn;
  CGCN_REQUIRE(comm != nullptr && ((send != nullptr && recv != nullptr) || bytes_per_rank == 0), "cgcn_comm_allgather: null argument");
  if (bytes_per_rank == 0) return CGCN_OK;
  NcclApi* api = nullptr;
  CGCN_TRY(need_nccl(&api));
  return nccl_status(api, api->AllGather(send, recv, bytes_per_rank, kNcclInt8, comm->nccl, static_cast<cudaStream_t>(stream)),
                     "ncclAllGather");
}
