// zfb_comm.cuh -- the one exchange step of the hot path, in the C ABI: metric partials (and histogram counters)
// combined across the GPUs of a job over NCCL / NVLink.
//
// The reference does this with blocking MPI_Allreduce calls on MPI_COMM_WORLD, three per metric object
// (absoluteError.hpp:80-92, relativeError.hpp:93-105, meansquareError.hpp:70-71, psnrError.hpp:75-83,
// minmaxMetric.hpp:77-81,122).  Here ONE ncclAllGather moves every rank's 64-byte zfb_partials and a one-block kernel
// combines them in rank order on every GPU: one collective on the critical path instead of two all-reduces plus
// packing kernels, and a result that does not depend on the reduction tree (bitwise equal on all ranks and to
// zfb_partials_combine on the host).
//
// NCCL is resolved at run time (dlopen): libzfb.so has no link-time dependency on it, and a process that already
// carries an NCCL (e.g. PyTorch's) gets that very library.  The communicator is the context's own
// (zfb_comm_unique_id on rank 0 -> broadcast the 128 bytes by any means, MPI_Bcast in CBench -> zfb_comm_init on every
// rank) or one the caller already has (zfb_comm_adopt; it must come from the same NCCL library).
#pragma once
#include <dlfcn.h>
#include <nccl.h>

namespace zfb {

struct nccl_api {
  void* lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  bool ok = false;
  std::string why;
};

static nccl_api& nccl()
{
  static nccl_api a;
  static bool tried = false;
  if (tried) return a;
  tried = true;
  const char* names[] = {std::getenv("ZFB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* nme : names) {
    if (!nme || !*nme) continue;
    a.lib = dlopen(nme, RTLD_NOW | RTLD_GLOBAL);
    if (a.lib) break;
  }
  if (!a.lib) { a.why = "libnccl.so.2 not found (set ZFB_NCCL_LIB)"; return a; }
#define ZFB_NCCL_SYM(field, sym)                                            \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, sym));         \
  if (!a.field) { a.why = std::string("NCCL symbol missing: ") + sym; return a; }
  ZFB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  ZFB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  ZFB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  ZFB_NCCL_SYM(AllGather, "ncclAllGather")
  ZFB_NCCL_SYM(AllReduce, "ncclAllReduce")
  ZFB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef ZFB_NCCL_SYM
  a.ok = true;
  return a;
}

// rank-ordered combine of the gathered partials (same arithmetic, same order as zfb_partials_combine)
__global__ void combine_partials_kernel(const partials_dev* __restrict__ all, int nranks, partials_dev* __restrict__ out)
{
  if (threadIdx.x || blockIdx.x) return;
  partials_dev t = all[0];
  for (int r = 1; r < nranks; r++) {
    const partials_dev s = all[r];
    t.sum_sq += s.sum_sq; t.sum_abs += s.sum_abs; t.sum_rel += s.sum_rel;
    t.max_abs = fmax(t.max_abs, s.max_abs); t.max_rel = fmax(t.max_rel, s.max_rel);
    t.max_orig = fmax(t.max_orig, s.max_orig); t.neg_min_orig = fmax(t.neg_min_orig, s.neg_min_orig);
    t.n += s.n;
  }
  *out = t;
}

}  // namespace zfb
