// comm.cu -- the one collective of the path behind the C ABI: the all-reduce (sum, float64) of
// the root statistics of a root-parallel search (SURVEY.md section 8e / Appendix E
// bmcts_allreduce_root; mamcts merge_node_statistics, hooks at
// /root/reference/bark_mcts/models/behavior/mcts_statistics/
// mcts_neural_cost_constrained_statistic.hpp:157-163,219-222).
//
// NCCL is bound at run time (dlopen of libnccl.so.2): libbmcts.so has no link-time dependency
// on it, a process that already carries an NCCL (PyTorch's) shares that copy, and a host
// without NCCL still loads the library -- bmcts_comm_* then fail with BMCTS_ERR_COMM.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <string>

#include "bmcts_c.h"
#include "bmcts_planner_c.h"

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi* Api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, []() {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (api.lib) break;
    }
    if (!api.lib) {
      api.error = "libnccl.so.2 not found";
      return;
    }
    auto sym = [&](const char* name) {
      void* s = dlsym(api.lib, name);
      if (!s) api.error = std::string("NCCL symbol missing: ") + name;
      return s;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  });
  return api.error.empty() ? &api : nullptr;
}

}  // namespace

struct bmcts_comm {
  int device = 0, nranks = 1, rank = 0;
  ncclComm_t comm = nullptr;
  cudaStream_t stream = nullptr;
  double* d_buf = nullptr;
  size_t cap = 0;
  std::string err;
};

static_assert(sizeof(ncclUniqueId) == BMCTS_COMM_ID_BYTES, "bmcts_c.h mirrors ncclUniqueId");

extern "C" {

int bmcts_comm_unique_id(void* id) {
  NcclApi* n = Api();
  if (!n || !id) return BMCTS_ERR_COMM;
  ncclUniqueId u;
  if (n->GetUniqueId(&u) != ncclSuccess) return BMCTS_ERR_COMM;
  std::memcpy(id, &u, sizeof(u));
  return BMCTS_OK;
}

int bmcts_comm_create(int device, int nranks, int rank, const void* id, bmcts_comm** out) {
  if (!out) return BMCTS_ERR_INVALID_ARG;
  *out = nullptr;
  if (!id || nranks < 1 || rank < 0 || rank >= nranks) return BMCTS_ERR_INVALID_ARG;
  NcclApi* n = Api();
  if (!n) return BMCTS_ERR_COMM;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return BMCTS_ERR_NO_DEVICE;
  if (device < 0 || device >= ndev) return BMCTS_ERR_INVALID_ARG;
  bmcts_comm* c = new bmcts_comm();
  c->device = device;
  c->nranks = nranks;
  c->rank = rank;
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof(u));
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      n->CommInitRank(&c->comm, nranks, u, rank) != ncclSuccess) {
    bmcts_comm_destroy(c);
    return BMCTS_ERR_COMM;
  }
  *out = c;
  return BMCTS_OK;
}

void bmcts_comm_destroy(bmcts_comm* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  NcclApi* n = Api();
  if (c->comm && n) n->CommDestroy(c->comm);
  if (c->d_buf) cudaFree(c->d_buf);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int bmcts_comm_nranks(const bmcts_comm* c) { return c ? c->nranks : 0; }

// Appendix E bmcts_allreduce_root: in-place sum over the ranks of n host doubles (the flattened
// root statistics: per ego action visits, return sum, cost sums; iterations; nodes -- < 1 KB)
int bmcts_allreduce_root(bmcts_comm* c, double* stats, size_t n) {
  if (!c || (!stats && n)) return BMCTS_ERR_INVALID_ARG;
  if (n == 0) return BMCTS_OK;
  NcclApi* api = Api();
  if (!api) return BMCTS_ERR_COMM;
  auto cu = [&](cudaError_t e, const char* what) {
    if (e == cudaSuccess) return false;
    c->err = std::string(what) + ": " + cudaGetErrorString(e);
    return true;
  };
  if (cu(cudaSetDevice(c->device), "cudaSetDevice")) return BMCTS_ERR_CUDA;
  if (n > c->cap) {
    if (c->d_buf) cudaFree(c->d_buf);
    c->d_buf = nullptr;
    c->cap = 0;
    if (cu(cudaMalloc((void**)&c->d_buf, sizeof(double) * (n + 64)), "cudaMalloc")) return BMCTS_ERR_CUDA;
    c->cap = n + 64;
  }
  if (cu(cudaMemcpyAsync(c->d_buf, stats, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream),
         "cudaMemcpyAsync"))
    return BMCTS_ERR_CUDA;
  const ncclResult_t r = api->AllReduce(c->d_buf, c->d_buf, n, ncclDouble, ncclSum, c->comm, c->stream);
  if (r != ncclSuccess) {
    c->err = std::string("ncclAllReduce: ") + api->GetErrorString(r);
    return BMCTS_ERR_COMM;
  }
  if (cu(cudaMemcpyAsync(stats, c->d_buf, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream),
         "cudaMemcpyAsync"))
    return BMCTS_ERR_CUDA;
  if (cu(cudaStreamSynchronize(c->stream), "cudaStreamSynchronize")) return BMCTS_ERR_CUDA;
  return BMCTS_OK;
}

// root statistics of the planner's last search -> all-reduce -> decision from the merged statistics
int bmcts_planner_allreduce_root_stats(bmcts_planner* p, bmcts_comm* comm, bmcts_plan_result* out) {
  if (!p || !comm || !out) return BMCTS_ERR_INVALID_ARG;
  double stats[4 * BMCTS_MAX_EGO_ACTIONS + 2];
  const int n = bmcts_planner_root_stats(p, stats, (int)(sizeof(stats) / sizeof(stats[0])));
  if (n <= 0) return BMCTS_ERR_INVALID_ARG;
  const int st = bmcts_allreduce_root(comm, stats, (size_t)n);
  if (st != BMCTS_OK) return st;
  return bmcts_planner_decide_from_root_stats(p, stats, n, out);
}

const char* bmcts_comm_last_error(const bmcts_comm* c) { return c ? c->err.c_str() : "null communicator"; }

}  // extern "C"
