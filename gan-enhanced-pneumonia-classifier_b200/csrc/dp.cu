// Data-parallel gradient-bucket layer of libb200gan.so: one NCCL communicator per process (one process per GPU), a dedicated
// communication stream, and event fork / join against the caller's compute stream, so that the all-reduce of a gradient bucket
// over NVLink / NVSwitch overlaps the rest of the backward pass -- eagerly or inside a CUDA-graph capture (the event record /
// wait pair forks the communication stream into the capture, b200gan_dp_sync joins it back).
//
// The reference has no multi-GPU path; the semantics are those of section 8e: sum of the per-rank
// gradients (the 1/world factor rides on b200gan_adam's grad_scale), BatchNorm statistics stay local.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 the process already holds -- torch's bundled 2.28 -- else the system one),
// so the library carries no link-time dependency on it and single-GPU users never load it.  Only the five entry points below are
// used; their ABI has been stable across NCCL 2.x.
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

namespace b200gan {

namespace {

struct NcclUniqueId { char internal[128]; };           // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* NcclComm;
enum { kNcclSuccess = 0, kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclSum = 0 };

struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*CommAbort)(NcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  bool ok = false;
};

const NcclApi& nccl() {
  static const NcclApi api = [] {
    NcclApi a;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // the copy the process (torch) already loaded, if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return a;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(h, "ncclAllReduce"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.CommAbort = reinterpret_cast<decltype(a.CommAbort)>(dlsym(h, "ncclCommAbort"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(dlsym(h, "ncclGetVersion"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.AllReduce && a.CommDestroy && a.GetErrorString;
    return a;
  }();
  return api;
}

int nccl_fail(int rc, const char* what) {
  set_error("NCCL error %d (%s) at %s", rc, nccl().GetErrorString ? nccl().GetErrorString(rc) : "?", what);
  return B200GAN_ERR_NCCL;
}

}  // namespace
}  // namespace b200gan

using namespace b200gan;

struct b200gan_dp {
  NcclComm comm = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  int world = 1, rank = 0, device = 0;
  int64_t collectives = 0;
};

#define B200_NCCL(expr)                                              \
  do {                                                               \
    int r__ = (expr);                                                \
    if (r__ != kNcclSuccess) return nccl_fail(r__, #expr);           \
  } while (0)

extern "C" {

int b200gan_dp_unique_id(void* id_out) {
  B200_CHECK_ARG(id_out, "dp_unique_id: null pointer");
  if (!nccl().ok) { set_error("dp_unique_id: libnccl.so.2 could not be loaded"); return B200GAN_ERR_NCCL; }
  B200_NCCL(nccl().GetUniqueId(reinterpret_cast<NcclUniqueId*>(id_out)));
  return 0;
}

int b200gan_dp_init(const void* id, int32_t world, int32_t rank, b200gan_dp** out) {
  B200_CHECK_ARG(id && out && world >= 1 && rank >= 0 && rank < world, "dp_init: bad argument (world=%d rank=%d)", world, rank);
  if (!nccl().ok) { set_error("dp_init: libnccl.so.2 could not be loaded"); return B200GAN_ERR_NCCL; }
  b200gan_dp* dp = new b200gan_dp();
  dp->world = world; dp->rank = rank;
  cudaError_t e = cudaGetDevice(&dp->device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&dp->comm_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&dp->fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&dp->join, cudaEventDisableTiming);
  if (e != cudaSuccess) { b200gan_dp_destroy(dp); return cuda_fail(e, "dp_init: stream / event creation"); }
  NcclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  int rc = nccl().CommInitRank(&dp->comm, world, uid, rank);
  if (rc != kNcclSuccess) { dp->comm = nullptr; b200gan_dp_destroy(dp); return nccl_fail(rc, "ncclCommInitRank"); }
  *out = dp;
  return 0;
}

int b200gan_dp_allreduce_bucket(b200gan_dp* dp, float* grad, int64_t numel, void* stream) {
  B200_CHECK_ARG(dp && dp->comm && grad && numel > 0, "dp_allreduce_bucket: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  // fork: everything the compute stream has launched so far (the kernels that finished this bucket's gradients) precedes the
  // collective; the compute stream itself goes on with the backward pass
  B200_CUDA(cudaEventRecord(dp->fork, st));
  B200_CUDA(cudaStreamWaitEvent(dp->comm_stream, dp->fork, 0));
  B200_NCCL(nccl().AllReduce(grad, grad, (size_t)numel, kNcclFloat32, kNcclSum, dp->comm, dp->comm_stream));
  dp->collectives++;
  return 0;
}

int b200gan_dp_allreduce_f64(b200gan_dp* dp, double* buf, int64_t numel, void* stream) {
  B200_CHECK_ARG(dp && dp->comm && buf && numel > 0, "dp_allreduce_f64: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  // synchronised BatchNorm: the per-channel sums of one layer, needed by the very next kernel.  Issued on the communication stream like the
  // gradient buckets (ONE stream per communicator keeps NCCL's issue order trivially identical on all ranks), fenced on both sides.
  B200_CUDA(cudaEventRecord(dp->fork, st));
  B200_CUDA(cudaStreamWaitEvent(dp->comm_stream, dp->fork, 0));
  B200_NCCL(nccl().AllReduce(buf, buf, (size_t)numel, kNcclFloat64, kNcclSum, dp->comm, dp->comm_stream));
  B200_CUDA(cudaEventRecord(dp->join, dp->comm_stream));
  B200_CUDA(cudaStreamWaitEvent(st, dp->join, 0));
  dp->collectives++;
  return 0;
}

int b200gan_dp_sync(b200gan_dp* dp, void* stream) {
  B200_CHECK_ARG(dp, "dp_sync: null handle");
  // join: the compute stream waits for every collective issued so far
  B200_CUDA(cudaEventRecord(dp->join, dp->comm_stream));
  B200_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, dp->join, 0));
  return 0;
}

int64_t b200gan_dp_collectives(const b200gan_dp* dp) { return dp ? dp->collectives : 0; }

int b200gan_dp_destroy(b200gan_dp* dp) {
  if (!dp) return 0;
  if (dp->comm_stream) cudaStreamSynchronize(dp->comm_stream);
  int rc = 0;
  if (dp->comm && nccl().ok) {
    // ncclCommAbort, not ncclCommDestroy: destroy is collective-flavoured (it waits for the peers and for every CUDA graph that
    // captured the communicator to be released first; measured: a rank whose captured iteration graph is still alive hangs
    // there at interpreter exit).  All work was drained above, so aborting only frees the resources, in any order across ranks.
    const int r = nccl().CommAbort ? nccl().CommAbort(dp->comm) : nccl().CommDestroy(dp->comm);
    if (r != kNcclSuccess) rc = nccl_fail(r, "ncclCommAbort");
  }
  if (dp->fork) cudaEventDestroy(dp->fork);
  if (dp->join) cudaEventDestroy(dp->join);
  if (dp->comm_stream) cudaStreamDestroy(dp->comm_stream);
  delete dp;
  return rc;
}

}  // extern "C"
