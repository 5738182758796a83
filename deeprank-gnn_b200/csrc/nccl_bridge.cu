// drgnn_nccl_*: the path's only collective as a C-ABI entry point - ONE ncclAllReduce(sum, fp32)
// of the flat [gradients | loss] buffer per step, for hosts whose GPUs cannot map each other's memory (no CUDA
// IPC / P2P): there the in-kernel NVLink exchange of fused_step2.cuh / comm.cu is unavailable.
//
// libdrgnn.so does NOT link NCCL: the library is bound at run time with dlopen, so that the .so loads on a box
// without NCCL and, inside a PyTorch process, the wrappers use the very libnccl.so.2 torch has already loaded
// (one NCCL per process).  Search order: $DRGNN_NCCL_LIB, "libnccl.so.2" (an already-loaded copy wins), "libnccl.so".
// Only the five stable entry points below are used; their prototypes are restated here (nccl.h 2.x: ncclUniqueId
// = 128 opaque bytes passed BY VALUE to ncclCommInitRank, ncclFloat32 = 7, ncclSum = 0).
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace drgnn {
namespace {

struct NcclUniqueId { char internal[DRGNN_NCCL_ID_BYTES]; };
typedef int (*get_unique_id_fn)(NcclUniqueId*);
typedef int (*comm_init_rank_fn)(void**, int, NcclUniqueId, int);
typedef int (*all_reduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*comm_destroy_fn)(void*);
typedef const char* (*get_error_string_fn)(int);

struct NcclApi {
  void* handle = nullptr;
  get_unique_id_fn get_unique_id = nullptr;
  comm_init_rank_fn comm_init_rank = nullptr;
  all_reduce_fn all_reduce = nullptr;
  comm_destroy_fn comm_destroy = nullptr;
  get_error_string_fn get_error_string = nullptr;
  char why[256] = {0};
};

NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* cands[3] = {getenv("DRGNN_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (int i = 0; i < 3 && !a.handle; ++i) {
      if (!cands[i] || !cands[i][0]) continue;
      a.handle = dlopen(cands[i], RTLD_NOW | RTLD_GLOBAL);
      if (!a.handle) {
        const char* e = dlerror();
        snprintf(a.why, sizeof(a.why), "dlopen(%s): %s", cands[i], e ? e : "?");
      }
    }
    if (!a.handle) return;
    a.get_unique_id = (get_unique_id_fn)dlsym(a.handle, "ncclGetUniqueId");
    a.comm_init_rank = (comm_init_rank_fn)dlsym(a.handle, "ncclCommInitRank");
    a.all_reduce = (all_reduce_fn)dlsym(a.handle, "ncclAllReduce");
    a.comm_destroy = (comm_destroy_fn)dlsym(a.handle, "ncclCommDestroy");
    a.get_error_string = (get_error_string_fn)dlsym(a.handle, "ncclGetErrorString");
    if (!a.get_unique_id || !a.comm_init_rank || !a.all_reduce || !a.comm_destroy) {
      snprintf(a.why, sizeof(a.why), "the NCCL library lacks ncclGetUniqueId / ncclCommInitRank / ncclAllReduce / ncclCommDestroy");
      a.get_unique_id = nullptr;
    }
  });
  return a;
}

inline bool ready(const NcclApi& a) { return a.handle && a.get_unique_id; }
inline const char* text(const NcclApi& a, int rc) { return a.get_error_string ? a.get_error_string(rc) : "NCCL error"; }

}  // namespace
}  // namespace drgnn

#define DRGNN_NCCL_READY(a)                                                                                   \
  do {                                                                                                        \
    if (!drgnn::ready(a)) return drgnn::fail(DRGNN_ERR_INVALID, "NCCL is not available: %s", (a).why[0] ? (a).why : "not found"); \
  } while (0)
#define DRGNN_NCCL_CALL(a, expr, what)                                                                        \
  do {                                                                                                        \
    const int _rc = (expr);                                                                                   \
    if (_rc != 0) return drgnn::fail(DRGNN_ERR_CUDA, "%s failed: %s (ncclResult %d)", what, drgnn::text(a, _rc), _rc); \
  } while (0)

extern "C" int drgnn_nccl_available(void) { return drgnn::ready(drgnn::api()) ? 1 : 0; }

extern "C" int drgnn_nccl_unique_id(void* id128) {
  DRGNN_REQUIRE(id128 != nullptr, "nccl_unique_id: NULL");
  drgnn::NcclApi& a = drgnn::api();
  DRGNN_NCCL_READY(a);
  drgnn::NcclUniqueId id;
  memset(&id, 0, sizeof(id));
  DRGNN_NCCL_CALL(a, a.get_unique_id(&id), "ncclGetUniqueId");
  memcpy(id128, &id, sizeof(id));
  return DRGNN_OK;
}

extern "C" int drgnn_nccl_init(void** comm, int32_t world, int32_t rank, const void* id128) {
  DRGNN_REQUIRE(comm != nullptr && id128 != nullptr, "nccl_init: NULL");
  DRGNN_REQUIRE(world >= 1 && rank >= 0 && rank < world, "nccl_init: rank %d outside a world of %d", rank, world);
  drgnn::NcclApi& a = drgnn::api();
  DRGNN_NCCL_READY(a);
  drgnn::NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  void* c = nullptr;
  DRGNN_NCCL_CALL(a, a.comm_init_rank(&c, (int)world, id, (int)rank), "ncclCommInitRank");
  *comm = c;
  return DRGNN_OK;
}

extern "C" int drgnn_nccl_allreduce(void* comm, float* buf, int64_t count, void* stream) {
  DRGNN_REQUIRE(comm != nullptr, "nccl_allreduce: no communicator");
  DRGNN_REQUIRE(count >= 0 && (buf != nullptr || count == 0), "nccl_allreduce: bad buffer");
  if (count == 0) return DRGNN_OK;
  drgnn::NcclApi& a = drgnn::api();
  DRGNN_NCCL_READY(a);
  // in place, fp32 (ncclFloat32 = 7), sum (ncclSum = 0), stream-ordered like every other entry point
  DRGNN_NCCL_CALL(a, a.all_reduce(buf, buf, (size_t)count, 7, 0, comm, (cudaStream_t)stream), "ncclAllReduce");
  return DRGNN_OK;
}

extern "C" int drgnn_nccl_destroy(void* comm) {
  if (!comm) return DRGNN_OK;
  drgnn::NcclApi& a = drgnn::api();
  DRGNN_NCCL_READY(a);
  DRGNN_NCCL_CALL(a, a.comm_destroy(comm), "ncclCommDestroy");
  return DRGNN_OK;
}
