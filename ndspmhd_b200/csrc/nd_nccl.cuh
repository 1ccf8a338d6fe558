// Native transport of the slab decomposition: NCCL called by the library itself, on the context's stream.
//
// The reference has no communication layer at all (it is serial, docs/about.rst:12).  The host-callback transport (`nd_comm`,
// torch.distributed or MPI on the caller's side) stays as the portable route; this one removes the host from the data path: every
// collective is enqueued on the compute stream between the kernels that produce and consume its operands --
//   all-reduces of the few scalars that steer the host loop (hhmax, unconverged count, relink/error flags, stressmax, vsigmax, dt's):
//       ncclAllReduce on a 16-double device block, then one store to pinned host memory and one stream synchronise
//       (against a C -> Python ctypes callback, pinned staging, torch all_reduce and two synchronises per reduction before);
//   the halo payloads: grouped ncclSend / ncclRecv between slab neighbours straight from the pack buffers into the unpack buffers,
//       stream-ordered, no synchronise;
//   the byte counts of a halo exchange: one ncclAllGather of two int64 per rank.
// libnccl is opened at run time (dlopen: the copy already in the process if the host program -- e.g. torch -- has one), so single-GPU
// users carry no dependency on it.  The host supplies only the 128-byte unique id (ndspmhd_b200_nccl_unique_id on rank 0, broadcast by
// whatever the host program has: torch.distributed, MPI_Bcast from Fortran).
#pragma once
#include <dlfcn.h>
#include <cstdlib>
#include <string>

extern "C" {
typedef struct ncclComm *nd_ncclComm_t;
typedef struct { char internal[128]; } nd_ncclUniqueId;   // NCCL_UNIQUE_ID_BYTES = 128 in every NCCL 2.x
}
// enum values of nccl.h (stable across NCCL 2.x)
enum { ND_NCCL_CHAR = 0, ND_NCCL_INT64 = 4, ND_NCCL_FLOAT64 = 8 };
enum { ND_NCCL_SUM = 0, ND_NCCL_MAX = 2, ND_NCCL_MIN = 3 };

struct NcclApi {
  void *handle = nullptr;
  int (*GetUniqueId)(nd_ncclUniqueId *) = nullptr;
  int (*CommInitRank)(nd_ncclComm_t *, int, nd_ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(nd_ncclComm_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, nd_ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, nd_ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, nd_ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, nd_ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  std::string why;
};

inline NcclApi *nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  const char *names[] = {getenv("NDSPMHD_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    if (!nm || !*nm) continue;
    api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
    api.why = dlerror();
  }
  if (!api.handle) return nullptr;
  bool ok = true;
  auto sym = [&](const char *n) { void *p = dlsym(api.handle, n); if (!p) { ok = false; api.why = std::string("missing symbol ") + n; } return p; };
  api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
  api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
  api.Send = (decltype(api.Send))sym("ncclSend");
  api.Recv = (decltype(api.Recv))sym("ncclRecv");
  api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  if (!ok) { dlclose(api.handle); api.handle = nullptr; return nullptr; }
  return &api;
}
