// comm.h -- the one exchange of the path: the sum of the per-GPU films, done by the LIBRARY.
// Replaces the sequential addTile merge of prender (Rendering.hs:130-134, Image.hs:178-199): each GPU renders a share of
// the sample indices into a private film; film_sum = sum over ranks (ncclAllReduce over NVLink / NVSwitch).
//
// NCCL is bound at run time (dlopen "libnccl.so.2": the copy a torch host has already loaded, else the system one), so the
// library has no link-time dependency and single-GPU hosts never touch it. The reduction runs on its own stream: it waits
// for the render calls enqueued so far (event), reads `film`, writes `film_sum`, and the NEXT render call only waits for it
// right before its film kernels add to `film` again -- so the all-reduce overlaps the next slice's raygen / traversal /
// shading instead of sitting between two slices.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <string>

namespace bl {

// the subset of nccl.h this file needs (NCCL 2.x ABI: ncclUniqueId is 128 opaque bytes, ncclComm_t an opaque pointer)
struct NcclUniqueId { char internal[128]; };
typedef void *NcclComm;
enum { kNcclSuccess = 0, kNcclFloat32 = 7, kNcclSum = 0 };

struct NcclApi {
   void *lib = nullptr;
   int (*GetVersion)(int *) = nullptr;
   int (*GetUniqueId)(NcclUniqueId *) = nullptr;
   int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
   int (*CommDestroy)(NcclComm) = nullptr;
   int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
   int (*Reduce)(const void *, void *, size_t, int, int, int, NcclComm, cudaStream_t) = nullptr;
   int (*GroupStart)() = nullptr;
   int (*GroupEnd)() = nullptr;
   const char *(*GetErrorString)(int) = nullptr;
   std::string err;

   bool load() {
      if (lib) return true;
      const char *names[] = {"libnccl.so.2", "libnccl.so"};
      for (const char *n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
      if (!lib) { err = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : ""); return false; }
      bool ok = true;
      auto sym = [&](const char *n) { void *p = dlsym(lib, n); if (!p) { ok = false; err = std::string("NCCL symbol missing: ") + n; } return p; };
      GetVersion = (int (*)(int *))sym("ncclGetVersion");
      GetUniqueId = (int (*)(NcclUniqueId *))sym("ncclGetUniqueId");
      CommInitRank = (int (*)(NcclComm *, int, NcclUniqueId, int))sym("ncclCommInitRank");
      CommDestroy = (int (*)(NcclComm))sym("ncclCommDestroy");
      AllReduce = (int (*)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t))sym("ncclAllReduce");
      Reduce = (int (*)(const void *, void *, size_t, int, int, int, NcclComm, cudaStream_t))sym("ncclReduce");
      GroupStart = (int (*)())sym("ncclGroupStart");
      GroupEnd = (int (*)())sym("ncclGroupEnd");
      GetErrorString = (const char *(*)(int))sym("ncclGetErrorString");
      if (!ok) { dlclose(lib); lib = nullptr; }
      return ok;
   }
   std::string what(int rc) const { return std::string("NCCL: ") + (GetErrorString ? GetErrorString(rc) : "error") + " (" + std::to_string(rc) + ")"; }
};
inline NcclApi &ncclApi() { static NcclApi a; return a; }

// per-context communicator state (owned by CudaBackend)
struct FilmComm {
   NcclComm comm = nullptr;
   int rank = 0, nranks = 1;
   cudaStream_t stream = nullptr;      // the reduction's own stream
   cudaEvent_t rendered = nullptr;     // compute stream -> comm stream: everything rendered so far
   cudaEvent_t reduced = nullptr;      // comm stream -> compute stream: film has been read, film_sum is complete
   bool pending = false;               // a reduction has been enqueued whose `reduced` event the compute stream has not waited for
   uint64_t reductions = 0, bytes = 0;
};

}  // namespace bl
