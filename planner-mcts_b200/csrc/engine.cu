// engine.cu -- C ABI of libbmcts.so: engine handle, device buffers, launches.
//
// One engine = one CUDA stream on one GPU with its own grow-only device
// buffers.  Batch calls with BMCTS_MEM_HOST pointers stage through the engine's
// device buffers (cudaMemcpyAsync on the engine stream, one copy per array) and
// return when the results are in the caller's memory; calls with
// BMCTS_MEM_DEVICE pointers only enqueue the kernel.  There is no CPU
// fallback: without a CUDA device bmcts_create fails with BMCTS_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "bmcts_c.h"
#include "inner_rects.hpp"
#include "rollout_device.cuh"
#include "rollout_lane_kernel.cuh"
#include "rollout_lane_coop_kernel.cuh"

namespace {

struct DevBuf {
  void* ptr = nullptr;
  size_t cap = 0;
};

}  // namespace

struct LaneGeometry {  // cached launch geometry of a lane kernel variant
  int W = 0, blocks = 0;
  size_t smem = 0;
};

struct bmcts_engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
  bool timing = true;  // record the events around every launch (bmcts_set_kernel_timing)
  // completion flag of zero-copy waves (pinned; written by the kernel's last warp)
  unsigned int* h_done_flag = nullptr;
  unsigned int done_seq = 0;
  bool flag_pending = false;  // the launch in flight raises the flag
  bool timed = false;
  bmcts_config cfg;
  bmcts_scenario scn;
  bool has_scenario = false;
  bmcts::InnerRects rects{};
  std::map<unsigned long long, LaneGeometry> geometry;
  // packed staging of small host-memory batches: one pinned block each way
  void* h_pack_in = nullptr;
  void* h_pack_out = nullptr;
  size_t h_pack_in_cap = 0, h_pack_out_cap = 0;
  // bmcts_expand_rollout_batch_begin .. _end: the download that is in flight
  bool async_pending = false;
  struct PendingItem {
    void* dst;
    size_t off, bytes;
  };
  std::vector<PendingItem> async_items;
  DevBuf d_pack_in, d_pack_out;
  bmcts_config* d_cfg = nullptr;
  bmcts_scenario* d_scn = nullptr;
  float* d_dt_table = nullptr;
  bool use_dt_table = false;
  uint64_t launches = 0;
  std::string err;
  int max_smem_optin = 0;
  int sm_count = 0;
  unsigned int* d_counter = nullptr;  // work counter of the lane kernel
  int lane_T = 0, lane_W = 0;          // last launch geometry (bmcts_last_launch_geometry)
  bool lane_ahead = false;
  // copy / compute pipeline of large host-memory rollout batches: chunk i+1 is uploaded and
  // chunk i-1 downloaded while chunk i computes
  struct PipeSlot {
    cudaStream_t stream = nullptr;
    unsigned int* d_counter = nullptr;
    DevBuf pose, alive, hyp, depth, index, result, ret_agents, final_pose;
  };
  static const int kPipeSlots = 3;
  PipeSlot pipe[kPipeSlots];
  cudaEvent_t ev_pipe = nullptr;
  unsigned int* d_ready = nullptr;   // [0] chunks uploaded, [1] wait timed out
  unsigned int* h_ready = nullptr;   // pinned: h_ready[c] = c + 1
  int h_ready_cap = 0;
  // staging buffers for host-pointer calls
  DevBuf b_pose, b_alive, b_hyp, b_depth, b_action, b_draw, b_index;
  DevBuf b_next_pose, b_next_alive, b_reward, b_cost, b_flags, b_last_action;
  DevBuf b_result, b_ret_agents, b_final_pose, b_prob, b_lik_action;
};

namespace {

int fail(bmcts_engine* e, int st, const std::string& msg) {
  if (e) e->err = msg;
  return st;
}

int cuda_fail(bmcts_engine* e, cudaError_t ce, const char* what) {
  return fail(e, BMCTS_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(ce));
}

#define CU(call)                                       \
  do {                                                 \
    cudaError_t ce_ = (call);                          \
    if (ce_ != cudaSuccess) return cuda_fail(e, ce_, #call); \
  } while (0)

int ensure(bmcts_engine* e, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return BMCTS_OK;
  if (b.ptr) CU(cudaFree(b.ptr));
  b.ptr = nullptr;
  b.cap = 0;
  size_t cap = bytes + bytes / 4 + 256;
  CU(cudaMalloc(&b.ptr, cap));
  b.cap = cap;
  return BMCTS_OK;
}

void free_buf(DevBuf& b) {
  if (b.ptr) cudaFree(b.ptr);
  b.ptr = nullptr;
  b.cap = 0;
}

int upload_dt_table(bmcts_engine* e) {
  // MctsStateBase::calculate_prediction_time_span (mcts_state_base.hpp:238-241)
  e->use_dt_table = e->cfg.prediction_alpha != 0.0f;
  if (!e->use_dt_table) return BMCTS_OK;
  std::vector<float> t(BMCTS_MAX_DEPTH);
  for (int d = 0; d < BMCTS_MAX_DEPTH; ++d)
    t[d] = (float)((double)e->cfg.prediction_k * std::pow((double)d, (double)e->cfg.prediction_alpha));
  CU(cudaMemcpyAsync(e->d_dt_table, t.data(), sizeof(float) * BMCTS_MAX_DEPTH,
                     cudaMemcpyHostToDevice, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return BMCTS_OK;
}

// Launch geometry of the lane kernel (rollout_lane_kernel.cuh).  T lanes per rollout: 1 when
// the batch fills the GPU, more when it does not (an MCTS wave of a few hundred leaves) or
// when the per-rollout state of a large scenario would leave too few warps resident.
// W warps per block share one staged scenario; the grid is persistent (blocks/SM x SMs).
// AHEAD: the "ego ahead" / "macro actions ahead" variants for batches too small to fill the GPU
// -- one warp of rollouts plus one warp that computes the macro actions a pass ahead per block
// (rollout_lane_kernel.cuh, rollout_lane_coop_kernel.cuh).
template <int T, bool COOP, bool AHEAD>
int launch_lane_T(bmcts_engine* e, const bmcts::KernelArgs& a, int force_W, cudaStream_t stream,
                  unsigned int* counter) {
  auto kern = COOP ? bmcts::rollout_lane_coop_kernel<T, AHEAD> : bmcts::rollout_lane_kernel<T, AHEAD>;
  auto layout = [&](int W) {
    bmcts::LaneLayout l =
        COOP ? bmcts::lane_coop_layout(a.A, a.C, T, W) : bmcts::lane_layout(a.A, a.C, a.H, T, W);
    if (AHEAD)
      l.total = bmcts::smem_align16(l.total) +
                (COOP ? sizeof(bmcts::CoopMail) : sizeof(bmcts::EgoMail));
    return l;
  };
  if (AHEAD) force_W = 1;
  const int NC = 32 / T;
  const long need_warps = ((long)a.n + NC - 1) / NC;
  // the kernel may use all the shared memory the SM offers.  cudaFuncSetAttribute applies to the
  // CURRENT device only, so the opt-in is made once per (variant, device): an engine on a second
  // GPU of the same process needs its own (engines on several host threads share the flags)
  constexpr int kMaxDevices = 64;
  static std::once_flag attr_once[kMaxDevices];
  static cudaError_t attr_err[kMaxDevices];
  if (e->device < 0 || e->device >= kMaxDevices)
    return fail(e, BMCTS_ERR_INVALID_ARG, "device ordinal out of range");
  const int optin = e->max_smem_optin;
  std::call_once(attr_once[e->device], [&]() {
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
    if (ce == cudaSuccess)
      ce = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared);
    attr_err[e->device] = ce;
  });
  if (attr_err[e->device] != cudaSuccess)
    return cuda_fail(e, attr_err[e->device], "cudaFuncSetAttribute");
  // geometry per (variant, scenario size, batch class): occupancy queries are cached
  // (COOP and AHEAD have the two top bits: T = 32 is 1 << 61 in the T field)
  const unsigned long long key = ((unsigned long long)COOP << 63) |
                                 ((unsigned long long)AHEAD << 62) | ((unsigned long long)T << 56) |
                                 ((unsigned long long)a.A << 48) | ((unsigned long long)a.C << 44) |
                                 ((unsigned long long)a.H << 36) | ((unsigned long long)force_W << 28) |
                                 (unsigned long long)std::min<long>(need_warps, (1L << 28) - 1);
  LaneGeometry g;
  auto it = e->geometry.find(key);
  if (it != e->geometry.end()) {
    g = it->second;
  } else {
    long best_score = -1;
    for (int W = BMCTS_LANE_MAX_THREADS / 32; W >= 1; --W) {
      if (force_W > 0 && W != force_W) continue;
      const bmcts::LaneLayout lay = layout(W);
      if ((int)lay.total > e->max_smem_optin) continue;
      int blocks = 0;
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kern, (W + (AHEAD ? 1 : 0)) * 32,
                                                       lay.total));
      if (blocks < 1) continue;
      // a small batch is spread over the SMs first: never more warps than it can fill
      long resident = (long)blocks * W * e->sm_count;
      long useful = resident < need_warps ? resident : need_warps;
      // prefer the geometry that runs most warps at once, then the one that spreads them
      // over more SMs (smaller W)
      long score = useful * 64 - (useful == need_warps ? W : 0);
      if (score > best_score) {
        best_score = score;
        g.W = W;
        g.blocks = blocks;
        g.smem = lay.total;
      }
    }
    if (best_score < 0)
      return fail(e, BMCTS_ERR_UNSUPPORTED, "scenario needs more shared memory than the SM offers");
    if (e->geometry.size() > 4096) e->geometry.clear();
    e->geometry[key] = g;
  }
  const long need_blocks = ((long)a.n + (long)NC * g.W - 1) / ((long)NC * g.W);
  long grid = (long)g.blocks * e->sm_count;
  if (grid > need_blocks) grid = need_blocks;
  if (const char* s = std::getenv("BMCTS_LANE_GRID")) {  // tests: force the refill path
    const long cap = std::atol(s);
    if (cap > 0 && grid > cap) grid = cap;
  }
  // (no memset of the work counter: the kernel's last warp clears it, retire_warp)
  kern<<<(unsigned)grid, (g.W + (AHEAD ? 1 : 0)) * 32, g.smem, stream>>>(a, counter);
  CU(cudaGetLastError());
  e->launches++;
  e->lane_T = T;
  e->lane_W = g.W;
  e->lane_ahead = AHEAD;
  if (std::getenv("BMCTS_DEBUG"))
    std::fprintf(stderr,
                 "[bmcts] lane kernel coop=%d ahead=%d T=%d W=%d blocks/SM=%d grid=%ld smem=%zu n=%d\n",
                 (int)COOP, (int)AHEAD, T, g.W, g.blocks, grid, g.smem, a.n);
  return BMCTS_OK;
}

// Lanes per rollout: the smallest T whose per-warp state (32/T rollouts) leaves about 20
// warps resident per SM; more when the batch is too small to fill the GPU with one lane per
// rollout (an MCTS wave of a few hundred leaves).
template <bool COOP>
int launch_lane(bmcts_engine* e, const bmcts::KernelArgs& a, cudaStream_t stream,
                unsigned int* counter) {
  int T = 0, W = 0;
  if (const char* s = std::getenv("BMCTS_LANE_T")) T = std::atoi(s);
  if (const char* s = std::getenv("BMCTS_LANE_W")) W = std::atoi(s);
  if (T != 1 && T != 2 && T != 4 && T != 8 && T != 16 && T != 32) {
    const size_t budget = 11 * 1024 + 512;
    auto warp_bytes = [&](int t) {
      return COOP ? bmcts::lane_coop_layout(a.A, a.C, t, 1).warp_bytes
                  : bmcts::lane_layout(a.A, a.C, a.H, t, 1).warp_bytes;
    };
    T = 1;
    if (warp_bytes(1) > budget) T = 2;
    if (warp_bytes(2) > budget) T = 4;
    // a batch too small to fill the GPU (an MCTS wave) is latency bound: more lanes per
    // rollout shorten the serial chain of a step (up to one agent per lane)
    const long lanes_resident = (long)e->sm_count * 32 * 16;
    for (int t = 2; t <= 32; t *= 2)
      if (T < t && (long)a.n * t <= lanes_resident && a.A > t / 2) T = t;
  }
  // the "ahead" variant when the batch leaves most of the GPU idle (latency bound: an MCTS wave;
  // measured on c2 / c5: faster up to about a quarter of the lanes the GPU can hold, slower from
  // half of them); the cooperative kernel's needs a lane per agent.  BMCTS_LANE_AHEAD = 0 / 1
  // forces the choice (tests)
  bool ahead = (long)a.n * T * 2 <= (long)e->sm_count * 32 * 16;
  if (const char* s = std::getenv("BMCTS_LANE_AHEAD")) ahead = std::atoi(s) != 0;
  if (COOP && a.A > T) ahead = false;
  if (ahead) {
    switch (T) {
      case 32: return launch_lane_T<32, COOP, true>(e, a, W, stream, counter);
      case 16: return launch_lane_T<16, COOP, true>(e, a, W, stream, counter);
      case 8: return launch_lane_T<8, COOP, true>(e, a, W, stream, counter);
      case 4: return launch_lane_T<4, COOP, true>(e, a, W, stream, counter);
      case 2: return launch_lane_T<2, COOP, true>(e, a, W, stream, counter);
      default: return launch_lane_T<1, COOP, true>(e, a, W, stream, counter);
    }
  }
  switch (T) {
    case 32: return launch_lane_T<32, COOP, false>(e, a, W, stream, counter);
    case 16: return launch_lane_T<16, COOP, false>(e, a, W, stream, counter);
    case 8: return launch_lane_T<8, COOP, false>(e, a, W, stream, counter);
    case 4: return launch_lane_T<4, COOP, false>(e, a, W, stream, counter);
    case 2: return launch_lane_T<2, COOP, false>(e, a, W, stream, counter);
    default: return launch_lane_T<1, COOP, false>(e, a, W, stream, counter);
  }
}

int launch_on(bmcts_engine* e, const bmcts::KernelArgs& a, cudaStream_t stream,
              unsigned int* counter) {
  if (e->cfg.state_kind == BMCTS_STATE_COOPERATIVE) return launch_lane<true>(e, a, stream, counter);
  return launch_lane<false>(e, a, stream, counter);
}

int launch(bmcts_engine* e, const bmcts::KernelArgs& a) {
  if (a.n <= 0) return BMCTS_OK;
  if (!e->timing) {  // bmcts_set_kernel_timing(e, 0): two driver calls per launch less
    e->timed = false;
    return launch_on(e, a, e->stream, e->d_counter);
  }
  CU(cudaEventRecord(e->ev_start, e->stream));
  int st = launch_on(e, a, e->stream, e->d_counter);
  if (st) return st;
  CU(cudaEventRecord(e->ev_stop, e->stream));
  e->timed = true;
  return BMCTS_OK;
}

void base_args(bmcts_engine* e, bmcts::KernelArgs& a, int n, uint64_t seed, uint32_t stream) {
  std::memset(&a, 0, sizeof(a));
  a.scn = e->d_scn;
  a.cfg = e->d_cfg;
  a.dt_table = e->use_dt_table ? e->d_dt_table : nullptr;
  a.n = n;
  a.ns = n;  // one start state per element unless a state index is given
  a.A = e->scn.n_agents;
  a.C = e->scn.n_corridors;
  a.H = e->scn.n_hypotheses;
  a.key0 = (uint32_t)seed;
  a.key1 = (uint32_t)(seed >> 32) ^ stream;
}

// stage one host array into a device buffer
int h2d(bmcts_engine* e, DevBuf& b, const void* src, size_t bytes, const void** dev) {
  *dev = nullptr;
  if (!src) return BMCTS_OK;
  int st = ensure(e, b, bytes);
  if (st) return st;
  CU(cudaMemcpyAsync(b.ptr, src, bytes, cudaMemcpyHostToDevice, e->stream));
  *dev = b.ptr;
  return BMCTS_OK;
}

int out_buf(bmcts_engine* e, DevBuf& b, const void* host, size_t bytes, void** dev) {
  *dev = nullptr;
  if (!host) return BMCTS_OK;
  int st = ensure(e, b, bytes);
  if (st) return st;
  *dev = b.ptr;
  return BMCTS_OK;
}

int d2h(bmcts_engine* e, void* dst, const void* dev, size_t bytes) {
  if (!dst || !dev) return BMCTS_OK;
  CU(cudaMemcpyAsync(dst, dev, bytes, cudaMemcpyDeviceToHost, e->stream));
  return BMCTS_OK;
}

// chunk size of the copy / compute pipeline for a host-memory batch of n rollouts (0 = one
// shot): chunks keep the persistent grid full (>= 128 Ki rollouts) and there are at least two
int pipeline_chunk(int n) {
  long chunk = 262144;
  if (const char* s = std::getenv("BMCTS_PIPE_CHUNK")) chunk = std::atol(s);
  if (chunk <= 0 || n < 2 * chunk) return 0;
  const long k = (n + chunk - 1) / chunk;
  long c = (n + k - 1) / k;
  c = (c + 31) & ~31L;
  return (int)c;
}

// chunk size of the streamed upload (0 = off)
int stream_chunk(int n) {
  // a kernel that waits for concurrent copies cannot run under a profiler that serialises
  // GPU work (ncu replays kernels in isolation): fall back to the chunked multi-kernel path
  static const bool profiled = std::getenv("CUDA_INJECTION64_PATH") != nullptr ||
                               std::getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") != nullptr ||
                               std::getenv("NVTX_INJECTION64_PATH") != nullptr;
  if (profiled && !std::getenv("BMCTS_STREAM_CHUNK")) return 0;
  if (const char* s = std::getenv("BMCTS_STREAM_CHUNK")) {
    const long chunk = std::atol(s);
    if (chunk <= 0 || n < 4 * chunk) return 0;
    return (int)((chunk + 31) & ~31L);
  }
  if (n < 32768) return 0;  // small batches: one packed upload, kernel, one packed download
  long chunk = ((long)n / 8 + 31) & ~31L;  // about eight chunks, 8 Ki .. 64 Ki rollouts each
  if (chunk < 8192) chunk = 8192;
  if (chunk > 65536) chunk = 65536;
  return (int)chunk;
}

// Large host-memory rollout batch, streamed: ONE persistent kernel over the whole batch is
// launched as soon as the first chunk of the leaf arrays is on its way; the remaining chunks
// are uploaded by the copy engine while the kernel runs, each followed by a 4-byte counter
// update, and a lane group only starts rollout k once the chunk holding k has arrived
// (wait_chunk in rollout_device.cuh; its spin is bounded, so a stalled upload ends in an error,
// not a hang).  Results are downloaded after the kernel.
int rollout_host_streamed(bmcts_engine* e, const bmcts_leaf_batch* in, const bmcts_rollout_out* out,
                          const bmcts::KernelArgs& base) {
  const int n = in->n_leaves, A = e->scn.n_agents;
  const int chunk = stream_chunk(n);
  const int n_chunks = (n + chunk - 1) / chunk;
  const bool indexed = in->n_states > 0;
  const int ns = indexed ? in->n_states : n;
  int st;
  if ((st = ensure(e, e->b_pose, sizeof(float) * 4 * (size_t)ns * A))) return st;
  if ((st = ensure(e, e->b_alive, sizeof(uint32_t) * (size_t)ns))) return st;
  if ((st = ensure(e, e->b_result, sizeof(bmcts_rollout_result) * (size_t)n))) return st;
  if (in->hyp_id && (st = ensure(e, e->b_hyp, (size_t)n * A))) return st;
  if (in->depth && (st = ensure(e, e->b_depth, sizeof(uint16_t) * (size_t)ns))) return st;
  if (indexed && (st = ensure(e, e->b_index, sizeof(uint32_t) * (size_t)n))) return st;
  if (out->ret_agents && (st = ensure(e, e->b_ret_agents, sizeof(float) * (size_t)n * A))) return st;
  if (out->final_pose && (st = ensure(e, e->b_final_pose, sizeof(float) * 4 * (size_t)n * A)))
    return st;
  if (n_chunks > e->h_ready_cap) {
    if (e->h_ready) CU(cudaFreeHost(e->h_ready));
    e->h_ready = nullptr;
    CU(cudaHostAlloc((void**)&e->h_ready, sizeof(unsigned int) * (size_t)(n_chunks + 64),
                     cudaHostAllocDefault));
    e->h_ready_cap = n_chunks + 64;
    for (int c = 0; c < e->h_ready_cap; ++c) e->h_ready[c] = (unsigned int)(c + 1);
  }
  cudaStream_t ks = e->stream, cs = e->pipe[0].stream;
  CU(cudaMemsetAsync(e->d_ready, 0, 2 * sizeof(unsigned int), ks));
  CU(cudaEventRecord(e->ev_pipe, ks));
  CU(cudaStreamWaitEvent(cs, e->ev_pipe, 0));  // uploads start after the counter reset
  bmcts::KernelArgs a = base;
  a.pose = (const float4*)e->b_pose.ptr;
  a.alive = (const uint32_t*)e->b_alive.ptr;
  a.hyp = in->hyp_id ? (const uint8_t*)e->b_hyp.ptr : nullptr;
  a.depth = in->depth ? (const uint16_t*)e->b_depth.ptr : nullptr;
  a.state_index = indexed ? (const uint32_t*)e->b_index.ptr : nullptr;
  a.result = (bmcts_rollout_result*)e->b_result.ptr;
  // result records go straight into the caller's buffer when it is pinned host memory the
  // device can address (32-byte posted writes over PCIe while the kernel runs): no download
  bool zero_copy_result = false;
  if (!std::getenv("BMCTS_NO_ZERO_COPY")) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, out->result) == cudaSuccess &&
        at.type == cudaMemoryTypeHost && at.devicePointer != nullptr) {
      a.result = (bmcts_rollout_result*)at.devicePointer;
      zero_copy_result = true;
    } else {
      (void)cudaGetLastError();  // an unregistered (pageable) pointer is not an error here
    }
  }
  a.ret_agents = out->ret_agents ? (float*)e->b_ret_agents.ptr : nullptr;
  a.final_pose = out->final_pose ? (float4*)e->b_final_pose.ptr : nullptr;
  a.ready = e->d_ready;
  a.chunk = (unsigned int)chunk;
  if (indexed) {
    // unique-state form: the (small) state table goes first, on the copy stream, so the counter
    // of chunk 0 also says that the table has arrived; the chunks carry hypothesis ids + indices
    CU(cudaMemcpyAsync(e->b_pose.ptr, in->pose, sizeof(float) * 4 * (size_t)ns * A,
                       cudaMemcpyHostToDevice, cs));
    CU(cudaMemcpyAsync(e->b_alive.ptr, in->alive, sizeof(uint32_t) * (size_t)ns,
                       cudaMemcpyHostToDevice, cs));
    if (in->depth)
      CU(cudaMemcpyAsync(e->b_depth.ptr, in->depth, sizeof(uint16_t) * (size_t)ns,
                         cudaMemcpyHostToDevice, cs));
  }
  for (int c = 0; c < n_chunks; ++c) {
    const size_t k0 = (size_t)c * chunk;
    const size_t m = (n - k0 < (size_t)chunk) ? n - k0 : (size_t)chunk;
    // one strided (2-D) copy per array and chunk: A rows of m elements, row pitch n elements
    if (!indexed)
      CU(cudaMemcpy2DAsync((char*)e->b_pose.ptr + k0 * 16, (size_t)n * 16, in->pose + k0 * 4,
                           (size_t)n * 16, m * 16, A, cudaMemcpyHostToDevice, cs));
    if (in->hyp_id)
      CU(cudaMemcpy2DAsync((char*)e->b_hyp.ptr + k0, (size_t)n, in->hyp_id + k0, (size_t)n, m, A,
                           cudaMemcpyHostToDevice, cs));
    if (indexed) {
      CU(cudaMemcpyAsync((char*)e->b_index.ptr + k0 * 4, in->state_index + k0, m * 4,
                         cudaMemcpyHostToDevice, cs));
    } else {
      CU(cudaMemcpyAsync((char*)e->b_alive.ptr + k0 * 4, in->alive + k0, m * 4,
                         cudaMemcpyHostToDevice, cs));
      if (in->depth)
        CU(cudaMemcpyAsync((char*)e->b_depth.ptr + k0 * 2, in->depth + k0, m * 2,
                           cudaMemcpyHostToDevice, cs));
    }
    CU(cudaMemcpyAsync(e->d_ready, e->h_ready + c, sizeof(unsigned int), cudaMemcpyHostToDevice, cs));
    if (c == 0) {  // the kernel starts while the first chunk is in flight
      CU(cudaEventRecord(e->ev_start, ks));
      if ((st = launch_on(e, a, ks, e->d_counter))) return st;
      CU(cudaEventRecord(e->ev_stop, ks));
      e->timed = true;
    }
  }
  unsigned int flags[2] = {0, 0};
  if (!zero_copy_result)
    CU(cudaMemcpyAsync(out->result, a.result, sizeof(bmcts_rollout_result) * (size_t)n,
                       cudaMemcpyDeviceToHost, ks));
  if (out->ret_agents)
    CU(cudaMemcpyAsync(out->ret_agents, a.ret_agents, sizeof(float) * (size_t)n * A,
                       cudaMemcpyDeviceToHost, ks));
  if (out->final_pose)
    CU(cudaMemcpyAsync(out->final_pose, a.final_pose, sizeof(float) * 4 * (size_t)n * A,
                       cudaMemcpyDeviceToHost, ks));
  CU(cudaMemcpyAsync(flags, e->d_ready, sizeof(flags), cudaMemcpyDeviceToHost, ks));
  CU(cudaStreamSynchronize(ks));
  CU(cudaStreamSynchronize(cs));
  if (flags[1]) return fail(e, BMCTS_ERR_CUDA, "streamed upload stalled: a chunk never arrived");
  return BMCTS_OK;
}

// Large host-memory rollout batch: the batch is cut into chunks of rollouts; every chunk is
// uploaded (strided 2-D copies out of the agent-major host arrays), rolled out and downloaded
// on its own stream, three chunks in flight, so the PCIe copies of one chunk overlap the kernel
// of another.  Results land in the caller's buffers before the call returns.
int rollout_host_pipelined(bmcts_engine* e, const bmcts_leaf_batch* in, const bmcts_rollout_out* out,
                           const bmcts::KernelArgs& base) {
  const int n = in->n_leaves, A = e->scn.n_agents;
  const int chunk = pipeline_chunk(n);
  const bool indexed = in->n_states > 0;
  const int ns = indexed ? in->n_states : n;
  int st;
  if (indexed) {  // the state table is shared by every chunk: uploaded once, ahead of the slots
    if ((st = ensure(e, e->b_pose, sizeof(float) * 4 * (size_t)ns * A))) return st;
    if ((st = ensure(e, e->b_alive, sizeof(uint32_t) * (size_t)ns))) return st;
    if (in->depth && (st = ensure(e, e->b_depth, sizeof(uint16_t) * (size_t)ns))) return st;
    CU(cudaMemcpyAsync(e->b_pose.ptr, in->pose, sizeof(float) * 4 * (size_t)ns * A,
                       cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->b_alive.ptr, in->alive, sizeof(uint32_t) * (size_t)ns,
                       cudaMemcpyHostToDevice, e->stream));
    if (in->depth)
      CU(cudaMemcpyAsync(e->b_depth.ptr, in->depth, sizeof(uint16_t) * (size_t)ns,
                         cudaMemcpyHostToDevice, e->stream));
  }
  // the slots' streams start after everything already queued on the engine stream
  CU(cudaEventRecord(e->ev_pipe, e->stream));
  for (auto& sl : e->pipe) CU(cudaStreamWaitEvent(sl.stream, e->ev_pipe, 0));
  int idx = 0;
  for (int k0 = 0; k0 < n; k0 += chunk, ++idx) {
    const int m = (n - k0 < chunk) ? n - k0 : chunk;
    bmcts_engine::PipeSlot& sl = e->pipe[idx % bmcts_engine::kPipeSlots];
    if (!indexed) {
      if ((st = ensure(e, sl.pose, sizeof(float) * 4 * (size_t)chunk * A))) return st;
      if ((st = ensure(e, sl.alive, sizeof(uint32_t) * (size_t)chunk))) return st;
      if (in->depth && (st = ensure(e, sl.depth, sizeof(uint16_t) * (size_t)chunk))) return st;
    } else if ((st = ensure(e, sl.index, sizeof(uint32_t) * (size_t)chunk))) {
      return st;
    }
    if ((st = ensure(e, sl.result, sizeof(bmcts_rollout_result) * (size_t)chunk))) return st;
    if (in->hyp_id && (st = ensure(e, sl.hyp, (size_t)chunk * A))) return st;
    if (out->ret_agents && (st = ensure(e, sl.ret_agents, sizeof(float) * (size_t)chunk * A))) return st;
    if (out->final_pose && (st = ensure(e, sl.final_pose, sizeof(float) * 4 * (size_t)chunk * A)))
      return st;
    cudaStream_t s = sl.stream;
    // ensure() may have re-allocated: safe, the slot's previous chunk is complete in stream order
    if (!indexed) {
      CU(cudaMemcpy2DAsync(sl.pose.ptr, (size_t)m * 16, in->pose + (size_t)k0 * 4, (size_t)n * 16,
                           (size_t)m * 16, A, cudaMemcpyHostToDevice, s));
      CU(cudaMemcpyAsync(sl.alive.ptr, in->alive + k0, sizeof(uint32_t) * (size_t)m,
                         cudaMemcpyHostToDevice, s));
      if (in->depth)
        CU(cudaMemcpyAsync(sl.depth.ptr, in->depth + k0, sizeof(uint16_t) * (size_t)m,
                           cudaMemcpyHostToDevice, s));
    } else {
      CU(cudaMemcpyAsync(sl.index.ptr, in->state_index + k0, sizeof(uint32_t) * (size_t)m,
                         cudaMemcpyHostToDevice, s));
    }
    if (in->hyp_id)
      CU(cudaMemcpy2DAsync(sl.hyp.ptr, (size_t)m, in->hyp_id + k0, (size_t)n, (size_t)m, A,
                           cudaMemcpyHostToDevice, s));
    bmcts::KernelArgs a = base;
    a.n = m;
    a.first_rollout_id = base.first_rollout_id + (uint64_t)k0;
    if (indexed) {
      a.ns = ns;
      a.pose = (const float4*)e->b_pose.ptr;
      a.alive = (const uint32_t*)e->b_alive.ptr;
      a.depth = in->depth ? (const uint16_t*)e->b_depth.ptr : nullptr;
      a.state_index = (const uint32_t*)sl.index.ptr;
    } else {
      a.ns = m;
      a.pose = (const float4*)sl.pose.ptr;
      a.alive = (const uint32_t*)sl.alive.ptr;
      a.depth = in->depth ? (const uint16_t*)sl.depth.ptr : nullptr;
    }
    a.hyp = in->hyp_id ? (const uint8_t*)sl.hyp.ptr : nullptr;
    a.result = (bmcts_rollout_result*)sl.result.ptr;
    a.ret_agents = out->ret_agents ? (float*)sl.ret_agents.ptr : nullptr;
    a.final_pose = out->final_pose ? (float4*)sl.final_pose.ptr : nullptr;
    if ((st = launch_on(e, a, s, sl.d_counter))) return st;
    CU(cudaMemcpyAsync(out->result + k0, sl.result.ptr, sizeof(bmcts_rollout_result) * (size_t)m,
                       cudaMemcpyDeviceToHost, s));
    if (out->ret_agents)
      CU(cudaMemcpy2DAsync(out->ret_agents + k0, (size_t)n * 4, sl.ret_agents.ptr, (size_t)m * 4,
                           (size_t)m * 4, A, cudaMemcpyDeviceToHost, s));
    if (out->final_pose)
      CU(cudaMemcpy2DAsync(out->final_pose + (size_t)k0 * 4, (size_t)n * 16, sl.final_pose.ptr,
                           (size_t)m * 16, (size_t)m * 16, A, cudaMemcpyDeviceToHost, s));
  }
  for (auto& sl : e->pipe) CU(cudaStreamSynchronize(sl.stream));
  return BMCTS_OK;
}

// ---- packed staging of small host-memory batches -------------------------------------
// A planner wave is a dozen small arrays each way.  Copying them one by one costs a dozen
// driver calls (and, from pageable memory, a dozen synchronous staging copies); instead they
// are gathered into ONE pinned block, uploaded with one copy, and the outputs come back the
// same way: 1 upload + 1 kernel + 1 download + 1 synchronize per wave.
struct Pack {
  struct Item {
    const void* src;
    void* dst;
    size_t off, bytes;
  };
  std::vector<Item> items;
  size_t bytes = 0;
  bool rewritten = false;  // an output the kernel writes more than once is part of the block
  static constexpr size_t kNone = ~size_t(0);
  size_t add(const void* src, void* dst, size_t n) {
    if (!src && !dst) return kNone;
    const size_t off = bytes;
    items.push_back({src, dst, off, n});
    bytes += (n + 255) & ~size_t(255);
    return off;
  }
};

int ensure_pinned(bmcts_engine* e, void** p, size_t* cap, size_t bytes) {
  if (bytes <= *cap) return BMCTS_OK;
  if (*p) CU(cudaFreeHost(*p));
  *p = nullptr;
  *cap = 0;
  const size_t want = bytes + bytes / 2 + 4096;
  CU(cudaHostAlloc(p, want, cudaHostAllocDefault));
  *cap = want;
  return BMCTS_OK;
}

constexpr size_t kPackLimit = 48u << 20;  // larger host batches keep the per-array / streamed paths

// The smallest waves of a tree search (a few dozen leaves, some tens of kilobytes): the kernel
// reads the packed inputs and writes the packed outputs straight in pinned host memory and the
// two copy commands of the wave disappear.  Measured on a B200: 31 instead of 32 ms for a
// 2000-iteration Plan() in waves of 15 leaves; no gain from 256-leaf waves on, hence the limit.
constexpr size_t kZeroCopyLimit = 64u << 10;

// (not when final poses are asked for: the kernel rewrites them at every step, which is cheap in
// device memory and would be a stream of small PCIe writes in host memory)
bool pack_zero_copy(const Pack& pin, const Pack& pout) {
  return pin.bytes + pout.bytes <= kZeroCopyLimit && !pout.rewritten &&
         !std::getenv("BMCTS_NO_ZERO_COPY");
}

// Waits for the wave in flight.  A zero-copy wave raises a flag in pinned memory when its last
// warp is done (retire_warp): the host reads that word instead of calling the driver, whose lock
// the host threads of a multi-tree search share; the stream is only queried now and then, so that
// a failed launch ends in an error and not in a spin.
int wait_wave(bmcts_engine* e) {
  if (!e->flag_pending) {
    CU(cudaStreamSynchronize(e->stream));
    return BMCTS_OK;
  }
  e->flag_pending = false;
  volatile unsigned int* f = e->h_done_flag;
  const unsigned int want = e->done_seq;
  for (unsigned int spins = 1; *f != want; ++spins) {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
    if ((spins & 2047u) == 0) {
      const cudaError_t ce = cudaStreamQuery(e->stream);
      if (ce == cudaSuccess) break;  // finished: the results are in place whatever the flag says
      if (ce != cudaErrorNotReady) return cuda_fail(e, ce, "cudaStreamQuery");
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  return BMCTS_OK;
}

// arms the completion flag for a launch that works on the pinned blocks directly
void arm_flag(bmcts_engine* e, bmcts::KernelArgs& a, const Pack& pin, const Pack& pout) {
  if (!pack_zero_copy(pin, pout) || !e->h_done_flag || std::getenv("BMCTS_NO_DONE_FLAG")) return;
  a.done_flag = e->h_done_flag;
  a.done_value = ++e->done_seq;
  e->flag_pending = true;
}

// stages the packed inputs (and uploads them unless the wave is small enough for zero copy);
// returns the base pointers of the input and output blocks as the kernel sees them
int pack_upload(bmcts_engine* e, const Pack& pin, const Pack& pout, char** d_in, char** d_out) {
  int st;
  if ((st = ensure_pinned(e, &e->h_pack_in, &e->h_pack_in_cap, pin.bytes))) return st;
  if ((st = ensure_pinned(e, &e->h_pack_out, &e->h_pack_out_cap, pout.bytes))) return st;
  for (const Pack::Item& it : pin.items)
    std::memcpy((char*)e->h_pack_in + it.off, it.src, it.bytes);
  if (pack_zero_copy(pin, pout)) {  // pinned memory is device-addressable (unified addressing)
    *d_in = (char*)e->h_pack_in;
    *d_out = (char*)e->h_pack_out;
    return BMCTS_OK;
  }
  if ((st = ensure(e, e->d_pack_in, pin.bytes))) return st;
  if ((st = ensure(e, e->d_pack_out, pout.bytes))) return st;
  CU(cudaMemcpyAsync(e->d_pack_in.ptr, e->h_pack_in, pin.bytes, cudaMemcpyHostToDevice, e->stream));
  *d_in = (char*)e->d_pack_in.ptr;
  *d_out = (char*)e->d_pack_out.ptr;
  return BMCTS_OK;
}

int pack_download(bmcts_engine* e, const Pack& pin, const Pack& pout, bool async = false) {
  if (pout.bytes && !pack_zero_copy(pin, pout)) {
    CU(cudaMemcpyAsync(e->h_pack_out, e->d_pack_out.ptr, pout.bytes, cudaMemcpyDeviceToHost,
                       e->stream));
  }
  if (async) {  // the caller collects with pack_collect (bmcts_expand_rollout_batch_end)
    e->async_items.clear();
    for (const Pack::Item& it : pout.items) e->async_items.push_back({it.dst, it.off, it.bytes});
    e->async_pending = true;
    return BMCTS_OK;
  }
  int st = wait_wave(e);
  if (st) return st;
  for (const Pack::Item& it : pout.items)
    std::memcpy(it.dst, (char*)e->h_pack_out + it.off, it.bytes);
  return BMCTS_OK;
}

int pack_collect(bmcts_engine* e) {
  e->async_pending = false;
  CU(cudaSetDevice(e->device));
  int st = wait_wave(e);
  if (st) return st;
  for (const bmcts_engine::PendingItem& it : e->async_items)
    std::memcpy(it.dst, (char*)e->h_pack_out + it.off, it.bytes);
  return BMCTS_OK;
}

int check_ready(bmcts_engine* e) {
  if (!e) return BMCTS_ERR_INVALID_ARG;
  if (!e->has_scenario) return fail(e, BMCTS_ERR_NO_SCENARIO, "bmcts_set_scenario was not called");
  if (e->async_pending)
    return fail(e, BMCTS_ERR_INVALID_ARG, "bmcts_expand_rollout_batch_end has not been called");
  CU(cudaSetDevice(e->device));
  return BMCTS_OK;
}

}  // namespace

extern "C" {

int bmcts_config_validate(const bmcts_config* c);

int bmcts_create(int device, const bmcts_config* cfg, bmcts_engine** out) {
  if (!out) return BMCTS_ERR_INVALID_ARG;
  *out = nullptr;
  // Root-parallel planners drive one engine (one stream) per tree.  With the default of 8
  // hardware connections streams share work queues and serialise each other; ask for the
  // maximum unless the application has decided.  Only effective if this is the first CUDA call
  // of the process -- an application that initialises CUDA earlier sets the variable itself.
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev <= 0) return BMCTS_ERR_NO_DEVICE;
  if (device < 0 || device >= ndev) return BMCTS_ERR_INVALID_ARG;
  bmcts_config c;
  if (cfg)
    c = *cfg;
  else
    bmcts_config_default(&c);
  if (bmcts_config_validate(&c) != BMCTS_OK) return BMCTS_ERR_INVALID_ARG;
  bmcts_engine* e = new bmcts_engine();
  e->device = device;
  e->cfg = c;
  auto bail = [&](cudaError_t err) {
    (void)err;
    bmcts_destroy(e);
    return BMCTS_ERR_CUDA;
  };
  if ((ce = cudaSetDevice(device)) != cudaSuccess) return bail(ce);
  if ((ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)) != cudaSuccess)
    return bail(ce);
  if ((ce = cudaEventCreate(&e->ev_start)) != cudaSuccess) return bail(ce);
  if ((ce = cudaEventCreate(&e->ev_stop)) != cudaSuccess) return bail(ce);
  if ((ce = cudaMalloc(&e->d_cfg, sizeof(bmcts_config))) != cudaSuccess) return bail(ce);
  // the scenario is followed by the derived inner rectangles of the road polygon
  if ((ce = cudaMalloc(&e->d_scn, sizeof(bmcts_scenario) + sizeof(bmcts::InnerRects))) != cudaSuccess)
    return bail(ce);
  if ((ce = cudaMalloc(&e->d_dt_table, sizeof(float) * BMCTS_MAX_DEPTH)) != cudaSuccess)
    return bail(ce);
  cudaDeviceGetAttribute(&e->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, device);
  if ((ce = cudaHostAlloc((void**)&e->h_done_flag, 64, cudaHostAllocDefault)) != cudaSuccess)
    return bail(ce);
  *e->h_done_flag = 0;
  if ((ce = cudaMalloc(&e->d_counter, 2 * sizeof(unsigned int))) != cudaSuccess) return bail(ce);
  if ((ce = cudaMemset(e->d_counter, 0, 2 * sizeof(unsigned int))) != cudaSuccess) return bail(ce);
  if ((ce = cudaEventCreateWithFlags(&e->ev_pipe, cudaEventDisableTiming)) != cudaSuccess)
    return bail(ce);
  if ((ce = cudaMalloc(&e->d_ready, 2 * sizeof(unsigned int))) != cudaSuccess) return bail(ce);
  for (auto& sl : e->pipe) {
    if ((ce = cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking)) != cudaSuccess)
      return bail(ce);
    if ((ce = cudaMalloc(&sl.d_counter, 2 * sizeof(unsigned int))) != cudaSuccess) return bail(ce);
    if ((ce = cudaMemset(sl.d_counter, 0, 2 * sizeof(unsigned int))) != cudaSuccess) return bail(ce);
  }
  if ((ce = cudaMemcpy(e->d_cfg, &e->cfg, sizeof(bmcts_config), cudaMemcpyHostToDevice)) !=
      cudaSuccess)
    return bail(ce);
  if (upload_dt_table(e) != BMCTS_OK) {
    bmcts_destroy(e);
    return BMCTS_ERR_CUDA;
  }
  *out = e;
  return BMCTS_OK;
}

void bmcts_destroy(bmcts_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  DevBuf* bufs[] = {&e->b_pose,      &e->b_alive,      &e->b_hyp,        &e->b_depth,
                    &e->b_action,    &e->b_draw,       &e->b_next_pose,  &e->b_next_alive,
                    &e->b_reward,    &e->b_cost,       &e->b_flags,      &e->b_last_action,
                    &e->b_result,    &e->b_ret_agents, &e->b_final_pose, &e->b_prob,
                    &e->b_lik_action, &e->b_index};
  for (DevBuf* b : bufs) free_buf(*b);
  if (e->d_cfg) cudaFree(e->d_cfg);
  if (e->d_scn) cudaFree(e->d_scn);
  if (e->d_dt_table) cudaFree(e->d_dt_table);
  if (e->d_counter) cudaFree(e->d_counter);
  if (e->h_done_flag) cudaFreeHost(e->h_done_flag);
  for (auto& sl : e->pipe) {
    if (sl.stream) cudaStreamSynchronize(sl.stream);
    DevBuf* sb[] = {&sl.pose, &sl.alive, &sl.hyp, &sl.depth, &sl.index, &sl.result, &sl.ret_agents, &sl.final_pose};
    for (DevBuf* b : sb) free_buf(*b);
    if (sl.d_counter) cudaFree(sl.d_counter);
    if (sl.stream) cudaStreamDestroy(sl.stream);
  }
  if (e->ev_pipe) cudaEventDestroy(e->ev_pipe);
  if (e->d_ready) cudaFree(e->d_ready);
  if (e->h_pack_in) cudaFreeHost(e->h_pack_in);
  if (e->h_pack_out) cudaFreeHost(e->h_pack_out);
  free_buf(e->d_pack_in);
  free_buf(e->d_pack_out);
  if (e->h_ready) cudaFreeHost(e->h_ready);
  if (e->ev_start) cudaEventDestroy(e->ev_start);
  if (e->ev_stop) cudaEventDestroy(e->ev_stop);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

int bmcts_set_config(bmcts_engine* e, const bmcts_config* cfg) {
  if (!e || !cfg) return BMCTS_ERR_INVALID_ARG;
  if (bmcts_config_validate(cfg) != BMCTS_OK) return fail(e, BMCTS_ERR_INVALID_ARG, "bad config");
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  e->cfg = *cfg;
  e->geometry.clear();  // the state kind (kernel variant, shared-memory layout) may have changed
  CU(cudaMemcpyAsync(e->d_cfg, &e->cfg, sizeof(bmcts_config), cudaMemcpyHostToDevice, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return upload_dt_table(e);
}

int bmcts_set_scenario(bmcts_engine* e, const bmcts_scenario* scn) {
  if (!e || !scn) return BMCTS_ERR_INVALID_ARG;
  if (!scn->finalized)
    return fail(e, BMCTS_ERR_INVALID_ARG, "scenario not finalized (bmcts_scenario_finalize)");
  if (e->cfg.state_kind != BMCTS_STATE_COOPERATIVE && scn->n_agents > 1 && scn->n_hypotheses < 1)
    return fail(e, BMCTS_ERR_INVALID_ARG, "hypothesis states need at least one hypothesis");
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  // inner rectangles: recomputed only when the road polygon changes (a planner sets the same
  // map at every Plan())
  const bool same_road = e->has_scenario && e->scn.n_road_pts == scn->n_road_pts &&
                         e->scn.n_corridors == scn->n_corridors &&
                         std::memcmp(e->scn.road_x, scn->road_x, sizeof(scn->road_x)) == 0 &&
                         std::memcmp(e->scn.road_y, scn->road_y, sizeof(scn->road_y)) == 0 &&
                         std::memcmp(e->scn.corridors, scn->corridors, sizeof(scn->corridors)) == 0;
  e->scn = *scn;
  if (!same_road) e->rects = bmcts::find_inner_rects(e->scn);
  bmcts::InnerRects up = e->rects;
  if (std::getenv("BMCTS_DEBUG"))
    for (int r = 0; r < up.n; ++r)
      std::fprintf(stderr, "[bmcts] inner rect %d: x [%g, %g] y [%g, %g]\n", r, up.rect[r][0],
                   up.rect[r][1], up.rect[r][2], up.rect[r][3]);
  if (std::getenv("BMCTS_NO_RECTS")) {  // tests: the scan-only path
    up.n = 0;
    for (int c = 0; c < BMCTS_MAX_CORRIDORS; ++c)
      for (int k = 0; k < BMCTS_MAX_CORRIDOR_PTS; ++k) {
        up.arc_lo[c][k] = 1.0e30f;
        up.arc_hi[c][k] = -1.0e30f;
      }
  }
  for (int r = up.n; r < 2; ++r) {  // unused rectangles are empty
    up.rect[r][0] = up.rect[r][2] = 1.0e30f;
    up.rect[r][1] = up.rect[r][3] = -1.0e30f;
  }
  CU(cudaMemcpyAsync(e->d_scn, &e->scn, sizeof(bmcts_scenario), cudaMemcpyHostToDevice,
                     e->stream));
  CU(cudaMemcpyAsync(reinterpret_cast<char*>(e->d_scn) + sizeof(bmcts_scenario), &up,
                     sizeof(bmcts::InnerRects), cudaMemcpyHostToDevice, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  e->has_scenario = true;
  return BMCTS_OK;
}

int bmcts_rollout_batch(bmcts_engine* e, const bmcts_leaf_batch* in, const bmcts_rollout_out* out) {
  int st = check_ready(e);
  if (st) return st;
  if (!in || !out || !in->pose || !in->alive || !out->result || in->n_leaves < 0)
    return fail(e, BMCTS_ERR_INVALID_ARG, "rollout_batch: null argument");
  const int n = in->n_leaves, A = e->scn.n_agents;
  if (n == 0) return BMCTS_OK;
  if (e->cfg.state_kind != BMCTS_STATE_COOPERATIVE && A > 1 && !in->hyp_id)
    return fail(e, BMCTS_ERR_INVALID_ARG, "rollout_batch: hyp_id required");
  bmcts::KernelArgs a;
  base_args(e, a, n, in->seed, in->stream);
  a.do_rollout = 1;
  a.first_rollout_id = in->first_rollout_id;
  if (e->cfg.max_rollout_steps <= 0)
    return fail(e, BMCTS_ERR_INVALID_ARG, "rollout_batch: max_rollout_steps must be > 0");
  const bool indexed = in->n_states > 0;
  if (in->n_states < 0 || (indexed && !in->state_index) || (!indexed && in->state_index))
    return fail(e, BMCTS_ERR_INVALID_ARG, "rollout_batch: n_states and state_index go together");
  const int ns = indexed ? in->n_states : n;  // rows of the pose / alive / depth tables
  a.ns = ns;
  if (in->memory == BMCTS_MEM_DEVICE) {
    a.pose = (const float4*)in->pose;
    a.alive = in->alive;
    a.hyp = in->hyp_id;
    a.depth = in->depth;
    a.state_index = in->state_index;
    a.result = out->result;
    a.ret_agents = out->ret_agents;
    a.final_pose = (float4*)out->final_pose;
    return launch(e, a);
  }
  if (stream_chunk(n) > 0) return rollout_host_streamed(e, in, out, a);
  if (pipeline_chunk(n) > 0) return rollout_host_pipelined(e, in, out, a);
  {
    const size_t nA = (size_t)n * A;
    Pack pin, pout;
    const size_t i_pose = pin.add(in->pose, nullptr, sizeof(float) * 4 * (size_t)ns * A);
    const size_t i_alive = pin.add(in->alive, nullptr, sizeof(uint32_t) * (size_t)ns);
    const size_t i_hyp = pin.add(in->hyp_id, nullptr, nA);
    const size_t i_depth = pin.add(in->depth, nullptr, sizeof(uint16_t) * (size_t)ns);
    const size_t i_index = pin.add(in->state_index, nullptr, sizeof(uint32_t) * (size_t)n);
    const size_t o_res = pout.add(nullptr, out->result, sizeof(bmcts_rollout_result) * (size_t)n);
    const size_t o_ra = pout.add(nullptr, out->ret_agents, sizeof(float) * nA);
    const size_t o_fp = pout.add(nullptr, out->final_pose, sizeof(float) * 4 * nA);
    pout.rewritten = out->final_pose != nullptr;
    if (pin.bytes + pout.bytes <= kPackLimit) {
      char *di, *dout;
      if ((st = pack_upload(e, pin, pout, &di, &dout))) return st;
      auto ip = [&](size_t off) -> const void* { return off == Pack::kNone ? nullptr : di + off; };
      auto op = [&](size_t off) -> void* { return off == Pack::kNone ? nullptr : dout + off; };
      a.pose = (const float4*)ip(i_pose);
      a.alive = (const uint32_t*)ip(i_alive);
      a.hyp = (const uint8_t*)ip(i_hyp);
      a.depth = (const uint16_t*)ip(i_depth);
      a.state_index = (const uint32_t*)ip(i_index);
      a.result = (bmcts_rollout_result*)op(o_res);
      a.ret_agents = (float*)op(o_ra);
      a.final_pose = (float4*)op(o_fp);
      arm_flag(e, a, pin, pout);
      if ((st = launch(e, a))) {
        e->flag_pending = false;
        return st;
      }
      return pack_download(e, pin, pout);
    }
  }
  const void* d;
  void* o;
  if ((st = h2d(e, e->b_pose, in->pose, sizeof(float) * 4 * (size_t)ns * A, &d))) return st;
  a.pose = (const float4*)d;
  if ((st = h2d(e, e->b_alive, in->alive, sizeof(uint32_t) * (size_t)ns, &d))) return st;
  a.alive = (const uint32_t*)d;
  if ((st = h2d(e, e->b_hyp, in->hyp_id, (size_t)n * A, &d))) return st;
  a.hyp = (const uint8_t*)d;
  if ((st = h2d(e, e->b_depth, in->depth, sizeof(uint16_t) * (size_t)ns, &d))) return st;
  a.depth = (const uint16_t*)d;
  if ((st = h2d(e, e->b_index, in->state_index, sizeof(uint32_t) * (size_t)n, &d))) return st;
  a.state_index = (const uint32_t*)d;
  if ((st = out_buf(e, e->b_result, out->result, sizeof(bmcts_rollout_result) * (size_t)n, &o)))
    return st;
  a.result = (bmcts_rollout_result*)o;
  if ((st = out_buf(e, e->b_ret_agents, out->ret_agents, sizeof(float) * (size_t)n * A, &o)))
    return st;
  a.ret_agents = (float*)o;
  if ((st = out_buf(e, e->b_final_pose, out->final_pose, sizeof(float) * 4 * (size_t)n * A, &o)))
    return st;
  a.final_pose = (float4*)o;
  if ((st = launch(e, a))) return st;
  if ((st = d2h(e, out->result, a.result, sizeof(bmcts_rollout_result) * (size_t)n))) return st;
  if ((st = d2h(e, out->ret_agents, a.ret_agents, sizeof(float) * (size_t)n * A))) return st;
  if ((st = d2h(e, out->final_pose, a.final_pose, sizeof(float) * 4 * (size_t)n * A))) return st;
  CU(cudaStreamSynchronize(e->stream));
  return BMCTS_OK;
}

static int step_common(bmcts_engine* e, const bmcts_step_in* in, const bmcts_step_out* so,
                       const bmcts_rollout_out* ro, uint64_t first_rollout_id, bool async = false) {
  int st = check_ready(e);
  if (st) return st;
  if (!in || !in->pose || !in->alive || !in->action || in->n < 0)
    return fail(e, BMCTS_ERR_INVALID_ARG, "step_batch: null argument");
  if (ro && !ro->result) return fail(e, BMCTS_ERR_INVALID_ARG, "expand_rollout: result required");
  if (!ro && !so) return fail(e, BMCTS_ERR_INVALID_ARG, "step_batch: no outputs");
  const int n = in->n, A = e->scn.n_agents;
  if (n == 0) return BMCTS_OK;
  const bool coop = e->cfg.state_kind == BMCTS_STATE_COOPERATIVE;
  if (!coop && A > 1 && (!in->hyp_id || !in->draw_id))
    return fail(e, BMCTS_ERR_INVALID_ARG, "step_batch: hyp_id and draw_id required");
  bmcts::KernelArgs a;
  base_args(e, a, n, in->seed, in->stream);
  a.do_expand = 1;
  a.do_rollout = ro ? 1 : 0;
  a.first_rollout_id = first_rollout_id;
  bmcts_step_out none;
  std::memset(&none, 0, sizeof(none));
  if (!so) so = &none;
  if (in->memory == BMCTS_MEM_DEVICE) {
    a.pose = (const float4*)in->pose;
    a.alive = in->alive;
    a.hyp = in->hyp_id;
    a.depth = in->depth;
    a.action = in->action;
    a.draw_id = in->draw_id;
    a.next_pose = (float4*)so->next_pose;
    a.next_alive = so->next_alive;
    a.reward = so->reward;
    a.cost = so->cost;
    a.flags = so->flags;
    a.last_action = so->last_action;
    if (ro) {
      a.result = ro->result;
      a.ret_agents = ro->ret_agents;
      a.final_pose = (float4*)ro->final_pose;
    }
    return launch(e, a);
  }
  const size_t nA = (size_t)n * A;
  {
    Pack pin, pout;
    const size_t i_pose = pin.add(in->pose, nullptr, sizeof(float) * 4 * nA);
    const size_t i_alive = pin.add(in->alive, nullptr, sizeof(uint32_t) * (size_t)n);
    const size_t i_hyp = pin.add(in->hyp_id, nullptr, nA);
    const size_t i_depth = pin.add(in->depth, nullptr, sizeof(uint16_t) * (size_t)n);
    const size_t i_action = pin.add(in->action, nullptr, nA);
    const size_t i_draw = pin.add(in->draw_id, nullptr, sizeof(uint64_t) * nA);
    const size_t o_np = pout.add(nullptr, so->next_pose, sizeof(float) * 4 * nA);
    const size_t o_na = pout.add(nullptr, so->next_alive, sizeof(uint32_t) * (size_t)n);
    const size_t o_rw = pout.add(nullptr, so->reward, sizeof(float) * nA);
    const size_t o_cost = pout.add(nullptr, so->cost, sizeof(float) * 2 * (size_t)n);
    const size_t o_fl = pout.add(nullptr, so->flags, sizeof(uint32_t) * (size_t)n);
    const size_t o_la = pout.add(nullptr, so->last_action, sizeof(float) * nA);
    const size_t o_res = pout.add(nullptr, ro ? ro->result : nullptr, sizeof(bmcts_rollout_result) * (size_t)n);
    const size_t o_ra = pout.add(nullptr, ro ? ro->ret_agents : nullptr, sizeof(float) * nA);
    const size_t o_fp = pout.add(nullptr, ro ? ro->final_pose : nullptr, sizeof(float) * 4 * nA);
    pout.rewritten = ro && ro->final_pose != nullptr;
    if (pin.bytes + pout.bytes <= kPackLimit) {
      char *di, *dout;
      if ((st = pack_upload(e, pin, pout, &di, &dout))) return st;
      auto ip = [&](size_t off) -> const void* { return off == Pack::kNone ? nullptr : di + off; };
      auto op = [&](size_t off) -> void* { return off == Pack::kNone ? nullptr : dout + off; };
      a.pose = (const float4*)ip(i_pose);
      a.alive = (const uint32_t*)ip(i_alive);
      a.hyp = (const uint8_t*)ip(i_hyp);
      a.depth = (const uint16_t*)ip(i_depth);
      a.action = (const uint8_t*)ip(i_action);
      a.draw_id = (const uint64_t*)ip(i_draw);
      a.next_pose = (float4*)op(o_np);
      a.next_alive = (uint32_t*)op(o_na);
      a.reward = (float*)op(o_rw);
      a.cost = (float*)op(o_cost);
      a.flags = (uint32_t*)op(o_fl);
      a.last_action = (float*)op(o_la);
      a.result = (bmcts_rollout_result*)op(o_res);
      a.ret_agents = (float*)op(o_ra);
      a.final_pose = (float4*)op(o_fp);
      arm_flag(e, a, pin, pout);
      if ((st = launch(e, a))) {
        e->flag_pending = false;
        return st;
      }
      return pack_download(e, pin, pout, async);
    }
  }
  // (a batch too large for the packed block runs synchronously also when called through _begin)
  const void* d;
  void* o;
  if ((st = h2d(e, e->b_pose, in->pose, sizeof(float) * 4 * nA, &d))) return st;
  a.pose = (const float4*)d;
  if ((st = h2d(e, e->b_alive, in->alive, sizeof(uint32_t) * (size_t)n, &d))) return st;
  a.alive = (const uint32_t*)d;
  if ((st = h2d(e, e->b_hyp, in->hyp_id, nA, &d))) return st;
  a.hyp = (const uint8_t*)d;
  if ((st = h2d(e, e->b_depth, in->depth, sizeof(uint16_t) * (size_t)n, &d))) return st;
  a.depth = (const uint16_t*)d;
  if ((st = h2d(e, e->b_action, in->action, nA, &d))) return st;
  a.action = (const uint8_t*)d;
  if ((st = h2d(e, e->b_draw, in->draw_id, sizeof(uint64_t) * nA, &d))) return st;
  a.draw_id = (const uint64_t*)d;
  if ((st = out_buf(e, e->b_next_pose, so->next_pose, sizeof(float) * 4 * nA, &o))) return st;
  a.next_pose = (float4*)o;
  if ((st = out_buf(e, e->b_next_alive, so->next_alive, sizeof(uint32_t) * (size_t)n, &o)))
    return st;
  a.next_alive = (uint32_t*)o;
  if ((st = out_buf(e, e->b_reward, so->reward, sizeof(float) * nA, &o))) return st;
  a.reward = (float*)o;
  if ((st = out_buf(e, e->b_cost, so->cost, sizeof(float) * 2 * (size_t)n, &o))) return st;
  a.cost = (float*)o;
  if ((st = out_buf(e, e->b_flags, so->flags, sizeof(uint32_t) * (size_t)n, &o))) return st;
  a.flags = (uint32_t*)o;
  if ((st = out_buf(e, e->b_last_action, so->last_action, sizeof(float) * nA, &o))) return st;
  a.last_action = (float*)o;
  if (ro) {
    if ((st = out_buf(e, e->b_result, ro->result, sizeof(bmcts_rollout_result) * (size_t)n, &o)))
      return st;
    a.result = (bmcts_rollout_result*)o;
    if ((st = out_buf(e, e->b_ret_agents, ro->ret_agents, sizeof(float) * nA, &o))) return st;
    a.ret_agents = (float*)o;
    if ((st = out_buf(e, e->b_final_pose, ro->final_pose, sizeof(float) * 4 * nA, &o))) return st;
    a.final_pose = (float4*)o;
  }
  if ((st = launch(e, a))) return st;
  if ((st = d2h(e, so->next_pose, a.next_pose, sizeof(float) * 4 * nA))) return st;
  if ((st = d2h(e, so->next_alive, a.next_alive, sizeof(uint32_t) * (size_t)n))) return st;
  if ((st = d2h(e, so->reward, a.reward, sizeof(float) * nA))) return st;
  if ((st = d2h(e, so->cost, a.cost, sizeof(float) * 2 * (size_t)n))) return st;
  if ((st = d2h(e, so->flags, a.flags, sizeof(uint32_t) * (size_t)n))) return st;
  if ((st = d2h(e, so->last_action, a.last_action, sizeof(float) * nA))) return st;
  if (ro) {
    if ((st = d2h(e, ro->result, a.result, sizeof(bmcts_rollout_result) * (size_t)n))) return st;
    if ((st = d2h(e, ro->ret_agents, a.ret_agents, sizeof(float) * nA))) return st;
    if ((st = d2h(e, ro->final_pose, a.final_pose, sizeof(float) * 4 * nA))) return st;
  }
  CU(cudaStreamSynchronize(e->stream));
  return BMCTS_OK;
}

int bmcts_step_batch(bmcts_engine* e, const bmcts_step_in* in, const bmcts_step_out* out) {
  if (!out) return e ? fail(e, BMCTS_ERR_INVALID_ARG, "step_batch: out is null") : BMCTS_ERR_INVALID_ARG;
  return step_common(e, in, out, nullptr, 0);
}

int bmcts_expand_rollout_batch(bmcts_engine* e, const bmcts_step_in* in,
                               const bmcts_step_out* step_out, uint64_t first_rollout_id,
                               const bmcts_rollout_out* rollout_out) {
  if (!rollout_out)
    return e ? fail(e, BMCTS_ERR_INVALID_ARG, "expand_rollout: rollout_out is null")
             : BMCTS_ERR_INVALID_ARG;
  return step_common(e, in, step_out, rollout_out, first_rollout_id);
}

int bmcts_expand_rollout_batch_begin(bmcts_engine* e, const bmcts_step_in* in,
                                     const bmcts_step_out* step_out, uint64_t first_rollout_id,
                                     const bmcts_rollout_out* rollout_out) {
  if (!rollout_out)
    return e ? fail(e, BMCTS_ERR_INVALID_ARG, "expand_rollout: rollout_out is null")
             : BMCTS_ERR_INVALID_ARG;
  if (in && in->memory == BMCTS_MEM_DEVICE)
    return fail(e, BMCTS_ERR_INVALID_ARG,
                "expand_rollout_begin: device buffers are asynchronous already (bmcts_sync)");
  return step_common(e, in, step_out, rollout_out, first_rollout_id, true);
}

int bmcts_expand_rollout_batch_end(bmcts_engine* e) {
  if (!e) return BMCTS_ERR_INVALID_ARG;
  if (!e->async_pending) return BMCTS_OK;  // empty or oversized batch: _begin has done it all
  return pack_collect(e);
}

int bmcts_likelihood_batch(bmcts_engine* e, const bmcts_likelihood_in* in, float* prob) {
  int st = check_ready(e);
  if (st) return st;
  if (!in || !prob || !in->pose || !in->alive || !in->action || in->n_worlds < 0 ||
      in->n_samples < 1 || in->n_buckets < 1 || !(in->buckets_upper > in->buckets_lower))
    return fail(e, BMCTS_ERR_INVALID_ARG, "likelihood_batch: bad argument");
  const int W = in->n_worlds, A = e->scn.n_agents, H = e->scn.n_hypotheses;
  if (W == 0 || H == 0) return BMCTS_OK;
  if (W > 65535 * 32 || H > 65535 || A > 65535)
    return fail(e, BMCTS_ERR_INVALID_ARG, "likelihood_batch: too many worlds");
  bmcts::LikelihoodArgs a;
  std::memset(&a, 0, sizeof(a));
  a.scn = e->d_scn;
  a.n_worlds = W;
  a.A = A;
  a.C = e->scn.n_corridors;
  a.H = H;
  a.n_samples = in->n_samples;
  a.n_buckets = in->n_buckets;
  a.lower = in->buckets_lower;
  a.upper = in->buckets_upper;
  a.key0 = (uint32_t)in->seed;
  a.key1 = (uint32_t)(in->seed >> 32) ^ in->stream;
  const size_t nprob = (size_t)W * A * H;
  if (in->memory == BMCTS_MEM_DEVICE) {
    a.pose = (const float4*)in->pose;
    a.alive = in->alive;
    a.action = in->action;
    a.prob = prob;
  } else {
    const void* d;
    if ((st = h2d(e, e->b_pose, in->pose, sizeof(float) * 4 * (size_t)W * A, &d))) return st;
    a.pose = (const float4*)d;
    if ((st = h2d(e, e->b_alive, in->alive, sizeof(uint32_t) * (size_t)W, &d))) return st;
    a.alive = (const uint32_t*)d;
    if ((st = h2d(e, e->b_lik_action, in->action, sizeof(float) * (size_t)W * A, &d))) return st;
    a.action = (const float*)d;
    if ((st = ensure(e, e->b_prob, sizeof(float) * nprob))) return st;
    a.prob = (float*)e->b_prob.ptr;
  }
  CU(cudaEventRecord(e->ev_start, e->stream));
  // hypotheses per block: as many as leave at least ~6 blocks per SM (the belief update of one
  // Plan() is a single world: few blocks, each must stay short)
  int hb = bmcts::kLikelihoodHypsPerBlock;
  while (hb > 1 && (long)W * A * ((H + hb - 1) / hb) < 6L * e->sm_count) hb /= 2;
  a.hyps_per_block = hb;
  dim3 grid((unsigned)W, (unsigned)((H + hb - 1) / hb), (unsigned)A);
  bmcts::likelihood_kernel<<<grid, 256, 0, e->stream>>>(a);
  CU(cudaGetLastError());
  e->launches++;
  CU(cudaEventRecord(e->ev_stop, e->stream));
  e->timed = true;
  if (in->memory != BMCTS_MEM_DEVICE) {
    if ((st = d2h(e, prob, a.prob, sizeof(float) * nprob))) return st;
    CU(cudaStreamSynchronize(e->stream));
  }
  return BMCTS_OK;
}

int bmcts_math_probe(bmcts_engine* e, int op, const float* a, const float* b, int n, float* out) {
  if (!e || !a || !b || !out || n < 0 || op < 0 || op >= BMCTS_MATH_NUM_OPS)
    return e ? fail(e, BMCTS_ERR_INVALID_ARG, "math_probe: bad argument") : BMCTS_ERR_INVALID_ARG;
  if (n == 0) return BMCTS_OK;
  if (e->async_pending)  // the probe shares the packed staging buffers with a wave in flight
    return fail(e, BMCTS_ERR_INVALID_ARG, "bmcts_expand_rollout_batch_end has not been called");
  CU(cudaSetDevice(e->device));
  const size_t bytes = sizeof(float) * (size_t)n;
  int st;
  if ((st = ensure(e, e->d_pack_in, 2 * bytes))) return st;
  if ((st = ensure(e, e->d_pack_out, bytes))) return st;
  float* da = (float*)e->d_pack_in.ptr;
  float* db = da + n;
  CU(cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, e->stream));
  CU(cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, e->stream));
  bmcts::math_probe_kernel<<<(n + 255) / 256, 256, 0, e->stream>>>(op, da, db, n,
                                                                    (float*)e->d_pack_out.ptr);
  CU(cudaGetLastError());
  e->launches++;
  CU(cudaMemcpyAsync(out, e->d_pack_out.ptr, bytes, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return BMCTS_OK;
}

int bmcts_alloc_host(size_t bytes, void** out) {
  if (!out) return BMCTS_ERR_INVALID_ARG;
  *out = nullptr;
  cudaError_t ce = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
  return ce == cudaSuccess ? BMCTS_OK : BMCTS_ERR_CUDA;
}

void bmcts_free_host(void* p) {
  if (p) cudaFreeHost(p);
}

int bmcts_sync(bmcts_engine* e) {
  if (!e) return BMCTS_ERR_INVALID_ARG;
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  return BMCTS_OK;
}

int bmcts_last_kernel_ms(bmcts_engine* e, float* ms) {
  if (!e || !ms) return BMCTS_ERR_INVALID_ARG;
  if (!e->timed)
    return fail(e, BMCTS_ERR_INVALID_ARG,
                e->timing ? "no batch call has been timed yet"
                          : "kernel timing is off (bmcts_set_kernel_timing)");
  CU(cudaSetDevice(e->device));
  CU(cudaEventSynchronize(e->ev_stop));
  CU(cudaEventElapsedTime(ms, e->ev_start, e->ev_stop));
  return BMCTS_OK;
}

int bmcts_last_launch_geometry(const bmcts_engine* e, int* lanes_per_rollout, int* warps_per_block,
                               int* ahead) {
  if (!e) return BMCTS_ERR_INVALID_ARG;
  if (lanes_per_rollout) *lanes_per_rollout = e->lane_T;
  if (warps_per_block) *warps_per_block = e->lane_W;
  if (ahead) *ahead = e->lane_ahead ? 1 : 0;
  return BMCTS_OK;
}

int bmcts_set_kernel_timing(bmcts_engine* e, int on) {
  if (!e) return BMCTS_ERR_INVALID_ARG;
  e->timing = on != 0;
  if (!e->timing) e->timed = false;
  return BMCTS_OK;
}

uint64_t bmcts_launch_count(const bmcts_engine* e) { return e ? e->launches : 0; }

void* bmcts_stream(bmcts_engine* e) { return e ? (void*)e->stream : nullptr; }

const char* bmcts_last_error(const bmcts_engine* e) { return e ? e->err.c_str() : "null engine"; }

}  // extern "C"
