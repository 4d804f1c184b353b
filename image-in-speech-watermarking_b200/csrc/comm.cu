// The two collectives of the path behind the C ABI: the all-reduce of the additive statistics vector at
// the end of an evaluation, and the all-reduce of the flat gradient buffer of the data-parallel training step.  NCCL is
// resolved at run time from the process (the library torch already loaded, else libnccl.so.2 on the loader path), so
// libwmk.so carries no link-time dependency on a particular NCCL build and still loads on a box without NCCL - the
// entry points then fail with WMK_ERR_UNSUPPORTED.  The communicator is the caller's ncclComm_t (any NCCL host code can
// pass its own), or one created here from a 128-byte unique id the caller distributes.
#include <dlfcn.h>
#include <mutex>
#include <string.h>

#include "wmk_common.cuh"

namespace wmk {
namespace {

// the slice of nccl.h this file needs (ABI-stable since NCCL 2.0)
typedef void* nccl_comm_t;
struct nccl_unique_id { char internal[128]; };
enum { kNcclSum = 0, kNcclFloat32 = 7, kNcclFloat64 = 8 };
typedef int (*get_unique_id_fn)(nccl_unique_id*);
typedef int (*comm_init_rank_fn)(nccl_comm_t*, int, nccl_unique_id, int);
typedef int (*comm_destroy_fn)(nccl_comm_t);
typedef int (*all_reduce_fn)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef const char* (*get_error_string_fn)(int);

struct NcclApi {
  get_unique_id_fn get_unique_id = nullptr;
  comm_init_rank_fn comm_init_rank = nullptr;
  comm_destroy_fn comm_destroy = nullptr;
  all_reduce_fn all_reduce = nullptr;
  get_error_string_fn get_error_string = nullptr;
  bool ok = false;
};

const NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    auto sym = [&](const char* name) -> void* {
      void* s = dlsym(RTLD_DEFAULT, name);
      if (s) return s;
      if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);     // the copy already mapped by the host process, if any
      return h ? dlsym(h, name) : nullptr;
    };
    api.get_unique_id = (get_unique_id_fn)sym("ncclGetUniqueId");
    api.comm_init_rank = (comm_init_rank_fn)sym("ncclCommInitRank");
    api.comm_destroy = (comm_destroy_fn)sym("ncclCommDestroy");
    api.all_reduce = (all_reduce_fn)sym("ncclAllReduce");
    api.get_error_string = (get_error_string_fn)sym("ncclGetErrorString");
    api.ok = api.get_unique_id && api.comm_init_rank && api.comm_destroy && api.all_reduce;
  });
  return api;
}

int nccl_check(int rc, const char* what) {
  if (rc == 0) return 0;
  const NcclApi& a = nccl();
  set_error("%s failed: NCCL error %d (%s)", what, rc, a.get_error_string ? a.get_error_string(rc) : "?");
  return WMK_ERR_CUDA;
}

int require_nccl() {
  if (!nccl().ok) { set_error("NCCL not found in this process (libnccl.so.2)"); return WMK_ERR_UNSUPPORTED; }
  return 0;
}

}  // namespace
}  // namespace wmk

using namespace wmk;

extern "C" int wmk_comm_unique_id(void* id128) {
  WMK_REQUIRE(id128, "comm_unique_id: null id");
  WMK_TRY(require_nccl());
  nccl_unique_id id;
  WMK_TRY(nccl_check(nccl().get_unique_id(&id), "ncclGetUniqueId"));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

extern "C" int wmk_comm_create(const void* id128, int nranks, int rank, void** comm) {
  WMK_REQUIRE(id128 && comm && nranks >= 1 && rank >= 0 && rank < nranks, "comm_create: bad arguments (nranks=%d rank=%d)", nranks, rank);
  WMK_TRY(require_nccl());
  nccl_unique_id id;
  memcpy(&id, id128, sizeof(id));
  nccl_comm_t c = nullptr;
  WMK_TRY(nccl_check(nccl().comm_init_rank(&c, nranks, id, rank), "ncclCommInitRank"));
  *comm = c;
  return 0;
}

extern "C" int wmk_comm_destroy(void* comm) {
  if (!comm) return 0;
  WMK_TRY(require_nccl());
  return nccl_check(nccl().comm_destroy((nccl_comm_t)comm), "ncclCommDestroy");
}

extern "C" int wmk_stats_allreduce_f64(double* stats, int n, void* nccl_comm, void* stream) {
  WMK_REQUIRE(stats && n > 0 && nccl_comm, "stats_allreduce: bad arguments");
  WMK_TRY(require_nccl());
  return nccl_check(nccl().all_reduce(stats, stats, (size_t)n, kNcclFloat64, kNcclSum, (nccl_comm_t)nccl_comm, (cudaStream_t)stream),
                    "ncclAllReduce(float64)");
}

extern "C" int wmk_grad_allreduce_f32(float* grads, size_t n, void* nccl_comm, void* stream) {
  WMK_REQUIRE(grads && n > 0 && nccl_comm, "grad_allreduce: bad arguments");
  WMK_TRY(require_nccl());
  return nccl_check(nccl().all_reduce(grads, grads, n, kNcclFloat32, kNcclSum, (nccl_comm_t)nccl_comm, (cudaStream_t)stream),
                    "ncclAllReduce(float32)");
}
