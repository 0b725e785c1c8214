// NCCL collective of the hot path behind the C ABI: ONE all-gather of the per-rank spectra slabs before consensus
// (replaces the trip through the filesystem of cnmf.py:748-773 `combine`).
//
// libnccl is bound at run time (dlopen): the library has no link-time dependency on it, a single-GPU host never
// loads it, and inside a PyTorch process the already-loaded copy (torch bundles 2.28.9) is reused.  A host that
// owns an ncclComm_t passes it to cnmf_allgather_spectra directly; one that does not can bootstrap a communicator
// with cnmf_comm_unique_id / cnmf_comm_create (the 128-byte id travels over whatever channel the host has).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <string>

#include "engine.h"

namespace {

struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
typedef int (*PFN_GetUniqueId)(NcclUniqueId*);
typedef int (*PFN_CommInitRank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*PFN_CommDestroy)(NcclComm);
typedef int (*PFN_AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
typedef const char* (*PFN_GetErrorString)(int);

struct NcclApi {
  void* lib = nullptr;
  PFN_GetUniqueId get_unique_id = nullptr;
  PFN_CommInitRank comm_init_rank = nullptr;
  PFN_CommDestroy comm_destroy = nullptr;
  PFN_AllGather all_gather = nullptr;
  PFN_GetErrorString error_string = nullptr;
  std::string why;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {                       // a copy already mapped into the process wins
      api.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD);
      if (api.lib) break;
    }
    if (!api.lib && std::getenv("CNMF_NCCL_LIB")) api.lib = dlopen(std::getenv("CNMF_NCCL_LIB"), RTLD_NOW | RTLD_GLOBAL);
    for (const char* n : names) {
      if (api.lib) break;
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!api.lib) {
      api.why = std::string("libnccl.so.2 could not be loaded (") + (dlerror() ? dlerror() : "?") +
                "); set CNMF_NCCL_LIB to its path";
      return;
    }
    api.get_unique_id = reinterpret_cast<PFN_GetUniqueId>(dlsym(api.lib, "ncclGetUniqueId"));
    api.comm_init_rank = reinterpret_cast<PFN_CommInitRank>(dlsym(api.lib, "ncclCommInitRank"));
    api.comm_destroy = reinterpret_cast<PFN_CommDestroy>(dlsym(api.lib, "ncclCommDestroy"));
    api.all_gather = reinterpret_cast<PFN_AllGather>(dlsym(api.lib, "ncclAllGather"));
    api.error_string = reinterpret_cast<PFN_GetErrorString>(dlsym(api.lib, "ncclGetErrorString"));
    if (!api.get_unique_id || !api.comm_init_rank || !api.comm_destroy || !api.all_gather)
      api.why = "libnccl is loaded but lacks ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy / ncclAllGather";
  });
  return api;
}

int nccl_fail(const char* what, int rc) {
  NcclApi& a = nccl();
  cnmf::set_last_error(std::string(what) + " failed: " + (a.error_string ? a.error_string(rc) : "ncclResult " + std::to_string(rc)));
  return -2;
}

#define CNMF_NCCL_READY()                              \
  do {                                                 \
    if (!nccl().why.empty() || !nccl().lib) {          \
      cnmf::set_last_error(nccl().why);                \
      return -3;                                       \
    }                                                  \
  } while (0)

constexpr int NCCL_FLOAT32 = 7;     // ncclFloat32 in nccl.h (stable since NCCL 2.0)

}  // namespace

extern "C" {

int cnmf_comm_unique_id(char* id_out_128) {
  CNMF_REQUIRE(id_out_128, "comm_unique_id: NULL output");
  CNMF_NCCL_READY();
  NcclUniqueId id;
  const int rc = nccl().get_unique_id(&id);
  if (rc != 0) return nccl_fail("ncclGetUniqueId", rc);
  std::memcpy(id_out_128, id.internal, sizeof(id.internal));
  return 0;
}

int cnmf_comm_create(cnmf_handle_t h, const char* id_128, int rank, int world, void** comm_out) {
  CNMF_REQUIRE(h && id_128 && comm_out && world >= 1 && rank >= 0 && rank < world, "comm_create: bad arguments");
  CNMF_NCCL_READY();
  CNMF_CUDA_CHECK(cudaSetDevice(h->device));
  NcclUniqueId id;
  std::memcpy(id.internal, id_128, sizeof(id.internal));
  NcclComm comm = nullptr;
  const int rc = nccl().comm_init_rank(&comm, world, id, rank);
  if (rc != 0) return nccl_fail("ncclCommInitRank", rc);
  *comm_out = comm;
  return 0;
}

int cnmf_comm_destroy(void* comm) {
  if (!comm) return 0;
  CNMF_NCCL_READY();
  const int rc = nccl().comm_destroy(comm);
  return rc == 0 ? 0 : nccl_fail("ncclCommDestroy", rc);
}

int cnmf_allgather_spectra(void* nccl_comm, const float* local_dev, long long rows_per_rank, long long ld,
                           float* merged_dev, void* stream) {
  CNMF_REQUIRE(nccl_comm && local_dev && merged_dev && rows_per_rank > 0 && ld > 0, "allgather_spectra: bad arguments");
  CNMF_NCCL_READY();
  const int rc = nccl().all_gather(local_dev, merged_dev, (size_t)rows_per_rank * (size_t)ld, NCCL_FLOAT32, nccl_comm,
                                   reinterpret_cast<cudaStream_t>(stream));
  return rc == 0 ? 0 : nccl_fail("ncclAllGather", rc);
}

}  // extern "C"
