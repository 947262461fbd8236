// Data-parallel gradient exchange from inside the library: tacorl_dp_allreduce_{init,enqueue,wait}.
// Replaces Lightning's DDP-over-gloo (/root/reference/config/trainer/default.yaml:1-4, scripts/train.py:73-75) with
// ncclAllReduce over NVLink / NVSwitch on a communication stream the library owns: `enqueue` orders the collective after
// everything already queued on the caller's stream and returns at once, `wait` makes the caller's stream wait for every
// exchange enqueued so far -- so a slice of the flat gradient travels while the encoder backward still runs.
// NCCL is reached through dlopen (no link-time dependency): the copy already loaded by torch, or libnccl.so.2 from the
// loader path, or TACORL_NCCL_LIB.  Host-side code only; no kernel of ours runs here.
#include "common.cuh"
#include "internal.h"
#include "../../include/tacorl_b200.h"
#include <dlfcn.h>
#include <cstdlib>

namespace tacorl {

struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
typedef int (*FnGetUniqueId)(NcclUniqueId*);
typedef int (*FnCommInitRank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*FnAllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*FnCommDestroy)(NcclComm);
typedef const char* (*FnGetErrorString)(int);
constexpr int kNcclFloat = 7, kNcclBfloat16 = 9, kNcclSum = 0;

struct DpState {
  void* lib = nullptr;
  FnGetUniqueId get_id = nullptr;
  FnCommInitRank init_rank = nullptr;
  FnAllReduce all_reduce = nullptr;
  FnCommDestroy destroy = nullptr;
  FnGetErrorString err = nullptr;
  NcclComm comm = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  int rank = 0, world = 1;
  bool pending = false;
};
static DpState g_dp;

static int dp_load() {
  if (g_dp.lib) return 0;
  const char* cands[4] = {getenv("TACORL_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
  void* h = nullptr;
  for (int i = 0; i < 3 && !h; ++i) {
    if (!cands[i]) continue;
    h = dlopen(cands[i], RTLD_NOW | RTLD_NOLOAD);        // the copy torch already mapped, if any
    if (!h) h = dlopen(cands[i], RTLD_NOW | RTLD_GLOBAL);
  }
  TACORL_REQUIRE(h, "dp_allreduce: cannot load NCCL (libnccl.so.2; set TACORL_NCCL_LIB): %s", dlerror());
  g_dp.get_id = (FnGetUniqueId)dlsym(h, "ncclGetUniqueId");
  g_dp.init_rank = (FnCommInitRank)dlsym(h, "ncclCommInitRank");
  g_dp.all_reduce = (FnAllReduce)dlsym(h, "ncclAllReduce");
  g_dp.destroy = (FnCommDestroy)dlsym(h, "ncclCommDestroy");
  g_dp.err = (FnGetErrorString)dlsym(h, "ncclGetErrorString");
  TACORL_REQUIRE(g_dp.get_id && g_dp.init_rank && g_dp.all_reduce && g_dp.destroy, "dp_allreduce: NCCL symbols missing");
  g_dp.lib = h;
  return 0;
}

#define TACORL_CHECK_NCCL(expr)                                                                         \
  do {                                                                                                  \
    const int _r = (expr);                                                                              \
    if (_r != 0) {                                                                                      \
      set_last_error("%s:%d NCCL error %d: %s", __FILE__, __LINE__, _r, g_dp.err ? g_dp.err(_r) : "?"); \
      return -3;                                                                                        \
    }                                                                                                   \
  } while (0)

}  // namespace tacorl

using namespace tacorl;

extern "C" {

int tacorl_dp_unique_id(void* id128) {
  TACORL_REQUIRE(id128, "dp_unique_id: null pointer");
  int rc;
  if ((rc = dp_load())) return rc;
  TACORL_CHECK_NCCL(g_dp.get_id((NcclUniqueId*)id128));
  return 0;
}

int tacorl_dp_allreduce_init(const void* id128, int rank, int world) {
  TACORL_REQUIRE(id128 && world >= 1 && rank >= 0 && rank < world, "dp_allreduce_init: bad arguments");
  TACORL_REQUIRE(!g_dp.comm, "dp_allreduce_init: already initialised (call tacorl_dp_allreduce_destroy first)");
  int rc;
  if ((rc = dp_load())) return rc;
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  TACORL_CHECK_NCCL(g_dp.init_rank(&g_dp.comm, world, id, rank));
  // highest priority: when a collective kernel is launched while compute grids fill the GPU (the early Adam update, the
  // encoder backward), its channel CTAs take the next SMs that free up instead of queueing behind the pending compute CTAs
  int prio_lo = 0, prio_hi = 0;
  TACORL_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  TACORL_CHECK_CUDA(cudaStreamCreateWithPriority(&g_dp.stream, cudaStreamNonBlocking, prio_hi));
  TACORL_CHECK_CUDA(cudaEventCreateWithFlags(&g_dp.fork, cudaEventDisableTiming));
  TACORL_CHECK_CUDA(cudaEventCreateWithFlags(&g_dp.join, cudaEventDisableTiming));
  g_dp.rank = rank; g_dp.world = world; g_dp.pending = false;
  return 0;
}

int tacorl_dp_allreduce_enqueue(void* buf, long long n, int dtype, void* stream) {
  TACORL_REQUIRE(g_dp.comm, "dp_allreduce_enqueue: call tacorl_dp_allreduce_init first");
  TACORL_REQUIRE(buf && n >= 0 && (dtype == 0 || dtype == 1), "dp_allreduce_enqueue: bad arguments");
  if (n == 0 || g_dp.world == 1) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  TACORL_CHECK_CUDA(cudaEventRecord(g_dp.fork, st));                    // after everything queued on the caller's stream
  TACORL_CHECK_CUDA(cudaStreamWaitEvent(g_dp.stream, g_dp.fork, 0));
  TACORL_CHECK_NCCL(g_dp.all_reduce(buf, buf, (size_t)n, dtype == 1 ? kNcclBfloat16 : kNcclFloat, kNcclSum, g_dp.comm,
                                    g_dp.stream));
  g_dp.pending = true;
  return 0;
}

int tacorl_dp_allreduce_wait(void* stream) {
  if (!g_dp.comm || !g_dp.pending) return 0;     // (pending stays set: several streams may wait for the same exchanges)
  TACORL_CHECK_CUDA(cudaEventRecord(g_dp.join, g_dp.stream));
  TACORL_CHECK_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, g_dp.join, 0));
  return 0;
}

int tacorl_dp_allreduce_destroy(void) {
  if (!g_dp.comm) return 0;
  cudaStreamSynchronize(g_dp.stream);
  g_dp.destroy(g_dp.comm);
  cudaEventDestroy(g_dp.fork);
  cudaEventDestroy(g_dp.join);
  cudaStreamDestroy(g_dp.stream);
  g_dp.comm = nullptr; g_dp.stream = nullptr; g_dp.fork = g_dp.join = nullptr; g_dp.pending = false;
  return 0;
}

}  // extern "C"
