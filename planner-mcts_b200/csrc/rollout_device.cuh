// rollout_device.cuh -- device functions shared by the sm_100a kernels of the MCTS
// leaf-evaluation path (Philox, Frenet projection, rectangle / polygon predicates, IDM,
// flags -> reward / cost / terminal) and the hypothesis-likelihood kernel (K4).
//
// The rollout kernels themselves are rollout_lane_kernel.cuh (hypothesis / risk-constraint
// states) and rollout_lane_coop_kernel.cuh (cooperative state).  Every float operation
// mirrors oracle/bmcts_oracle.c operation by operation (compile with --fmad=false; FMAs only
// where bm_fma is written).  Paths cited are relative to
// /root/reference/bark_mcts/models/behavior/.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "bmcts_c.h"
#include "bmcts_math.h"

namespace bmcts {

constexpr float kBig = 1.0e30f;
constexpr float kGapEps = 1.0e-3f;
constexpr uint32_t kGoalFailBit = 1u << 16;  // scratch bit: a goal-polygon corner is outside

static_assert(sizeof(bmcts_scenario) % 16 == 0, "scenario is staged with 16-byte copies");
static_assert(sizeof(bmcts_config) % 4 == 0, "config is staged with 4-byte copies");

struct KernelArgs {
  // inputs (agent-major SoA, see bmcts_c.h)
  const float4* pose;
  const uint32_t* alive;
  const uint8_t* hyp;
  const uint16_t* depth;
  const uint8_t* action;     // expansion only
  const uint64_t* draw_id;   // expansion only
  // step outputs (expansion), each may be null
  float4* next_pose;
  uint32_t* next_alive;
  float* reward;
  float* cost;
  uint32_t* flags;
  float* last_action;
  // rollout outputs
  bmcts_rollout_result* result;
  float* ret_agents;
  float4* final_pose;
  // device copies of scenario / config / prediction-time table
  const bmcts_scenario* scn;
  const bmcts_config* cfg;
  const float* dt_table;  // null when prediction_alpha == 0
  int n;
  int A;
  int C;
  int H;
  int do_expand;
  int do_rollout;
  uint32_t key0, key1;
  uint64_t first_rollout_id;
  // streamed host batch (engine.cu rollout_host_streamed): rollout k may start once
  // ready[0] > k / chunk, i.e. its chunk of the leaf arrays has been uploaded; ready[1] is
  // set when a wait times out.  Null for every other call.
  unsigned int* ready;
  unsigned int chunk;
  // unique-state form of a leaf batch (bmcts_leaf_batch.state_index): pose / alive / depth are
  // tables of ns states and rollout k starts from state state_index[k]; null: ns == n, state k
  const uint32_t* state_index;
  int ns;
  // completion flag of a wave that runs on pinned host memory (engine.cu pack_zero_copy): the
  // last warp of the grid stores done_value there after its results, and the host waits on the
  // flag instead of making a driver call.  Null for every other call.
  unsigned int* done_flag;
  unsigned int done_value;
};

// start pose of agent a of rollout rid: row rid of the agent-major batch arrays, or -- unique-state
// form -- record state_index[rid] of the STATE-MAJOR table [ns][A] (one state is 16 A contiguous
// bytes: a gather in any order reads whole sectors)
__device__ __forceinline__ float4 start_pose(const KernelArgs& p, int a, uint32_t rid) {
  return p.state_index ? p.pose[(size_t)p.state_index[rid] * p.A + a] : p.pose[(size_t)a * p.n + rid];
}

__host__ __device__ inline size_t smem_align16(size_t x) { return (x + 15) & ~size_t(15); }

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

__device__ __forceinline__ uint4 draw(uint32_t k0, uint32_t k1, uint64_t id, uint32_t agent,
                                      uint32_t domain, uint32_t index) {
  return philox4x32_10(
      make_uint4((uint32_t)id, (uint32_t)(id >> 32), agent | (domain << 16), index), k0, k1);
}

// wait until the chunk that holds rollout `rid` has arrived (acquire load at system scope:
// the data is written by the copy engine, then the counter).  Gives up after two seconds
// instead of hanging the GPU if the upload stalled (e.g. a tool that serialises GPU work).
__device__ __forceinline__ bool wait_chunk(const KernelArgs& p, uint32_t rid) {
  if (!p.ready) return true;
  const unsigned int need = rid / p.chunk + 1u;
  unsigned long long t0 = 0;
  for (unsigned int spins = 0;; ++spins) {
    unsigned int have;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(have) : "l"(p.ready) : "memory");
    if (have >= need) return true;
    __nanosleep(100);
    if ((spins & 1023u) == 0u) {  // wall-clock bound: 2 s without the chunk = the upload stalled
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      if (now - t0 > 2000000000ull) break;
    }
  }
  atomicExch(p.ready + 1, 1u);
  return false;
}

// unroll factors as build-time macros (tools/sweep_unroll.sh): a pragma does not expand macros
#define BM_UNROLL_STR(x) #x
#define BM_UNROLL(n) _Pragma(BM_UNROLL_STR(unroll n))
#ifndef BMCTS_PROJECT_UNROLL
#define BMCTS_PROJECT_UNROLL 1
#endif

struct Proj {
  float s, d;
  int seg;
  bool in;
};

// Frenet projection on a corridor centre line (oracle_project).  Corridors have a handful
// of segments: the loop is kept rolled (an unrolled-by-4 body with its remainder loops
// costs more than it saves at trip counts of 1-3) and the best-so-far update is selects.
__device__ __forceinline__ Proj project(const bmcts_corridor& c, float x, float y) {
  const int nseg = c.n_pts - 1;
  float best_d2 = kBig, best_u = 0.0f, best_raw = 0.0f, best_cross = 0.0f;
  int best = 0;
BM_UNROLL(BMCTS_PROJECT_UNROLL)
  for (int k = 0; k < nseg; ++k) {
    const float tx = c.tx[k], ty = c.ty[k];
    const float rx = x - c.px[k];
    const float ry = y - c.py[k];
    const float u = bm_fma(rx, tx, ry * ty);
    const float uc = bm_clamp(u, 0.0f, c.seg_len[k]);
    const float ex = bm_fma(-uc, tx, rx);
    const float ey = bm_fma(-uc, ty, ry);
    const float d2 = bm_fma(ex, ex, ey * ey);
    const float cross = bm_fma(tx, ry, -(ty * rx));
    const bool better = d2 < best_d2;
    best_d2 = better ? d2 : best_d2;
    best = better ? k : best;
    best_u = better ? uc : best_u;
    best_raw = better ? u : best_raw;
    best_cross = better ? cross : best_cross;
  }
  Proj p;
  p.s = c.s0[best] + best_u;
  p.d = best_cross;
  p.seg = best;
  p.in = !((best == 0 && best_raw < 0.0f) || (best == nseg - 1 && best_raw > c.seg_len[best]));
  return p;
}

// even-odd rule, division free (oracle point_in_polygon)
__device__ __forceinline__ bool point_in_polygon(const float* px, const float* py, int n, float x,
                                                 float y) {
  int inside = 0;
  float xj = px[n - 1], yj = py[n - 1];
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
    const float xi = px[i], yi = py[i];
    if ((yi > y) != (yj > y)) {
      const float lhs = (x - xi) * (yj - yi);
      const float rhs = (xj - xi) * (y - yi);
      const int cross = (yj > yi) ? (lhs < rhs) : (lhs > rhs);
      inside ^= cross;
    }
    xj = xi;
    yj = yi;
  }
  return inside != 0;
}

struct Box {
  float cx, cy, c, s, hl, hw;
};

__device__ __forceinline__ Box agent_box(const bmcts_scenario& scn, int a, float x, float y,
                                         float c, float s, float lat, float lon) {
  float front = (scn.agent_length[a] - scn.agent_rear[a]) + lon;
  float back = scn.agent_rear[a] + lon;
  Box b;
  b.hl = 0.5f * (front + back);
  float off = 0.5f * (front - back);
  b.cx = bm_fma(off, c, x);
  b.cy = bm_fma(off, s, y);
  b.c = c;
  b.s = s;
  b.hw = bm_fma(0.5f, scn.agent_width[a], lat);
  return b;
}

__device__ __forceinline__ bool boxes_overlap(const Box& A, const Box& B) {
  float tx = B.cx - A.cx;
  float ty = B.cy - A.cy;
  float cd = bm_fma(A.c, B.c, A.s * B.s);
  float sd = bm_fma(A.c, B.s, -(A.s * B.c));
  float acd = bm_abs(cd), asd = bm_abs(sd);
  float p0 = bm_abs(bm_fma(tx, A.c, ty * A.s));
  bool sep = p0 >= A.hl + bm_fma(B.hl, acd, B.hw * asd);
  float p1 = bm_abs(bm_fma(ty, A.c, -(tx * A.s)));
  sep = sep || (p1 >= A.hw + bm_fma(B.hl, asd, B.hw * acd));
  float p2 = bm_abs(bm_fma(tx, B.c, ty * B.s));
  sep = sep || (p2 >= B.hl + bm_fma(A.hl, acd, A.hw * asd));
  float p3 = bm_abs(bm_fma(ty, B.c, -(tx * B.s)));
  sep = sep || (p3 >= B.hw + bm_fma(A.hl, asd, A.hw * acd));
  return !sep;
}

__device__ __forceinline__ void box_corner(const Box& b, int k, float& x, float& y) {
  float lx = (k == 0 || k == 1) ? b.hl : -b.hl;
  float ly = (k == 0 || k == 3) ? b.hw : -b.hw;
  x = b.cx + (lx * b.c - ly * b.s);
  y = b.cy + (lx * b.s + ly * b.c);
}

__device__ __forceinline__ bool corner_in_polygon(const Box& b, int k, const float* px,
                                                  const float* py, int n) {
  float x, y;
  box_corner(b, k, x, y);
  return point_in_polygon(px, py, n, x, y);
}

// Rizaldi et al. longitudinal safe distance (oracle safe_dist_longitudinal)
__device__ __forceinline__ bool safe_dist_longitudinal(float v_f, float v_r, float dist, float a_r,
                                                       float a_f, float delta) {
  if (dist < 0.0f) return false;
  float t_stop_f = bm_div(-v_f, a_f);
  float v_f_star = (delta <= t_stop_f) ? bm_fma(a_f, delta, v_f) : 0.0f;
  float t_stop_f_star = bm_div(-v_f_star, a_f);  // the front vehicle's own stopping time
  float t_stop_r = bm_div(-v_r, a_r);
  float sd0 = v_r * delta - bm_div(v_r * v_r, 2.0f * a_r);
  float sd1 = sd0 + bm_div(v_f * v_f, 2.0f * a_f);
  float front_react = v_f * delta + (0.5f * a_f) * (delta * delta);
  float sd3 = sd0 - front_react;
  if (dist > sd0 || (delta <= t_stop_f && dist > sd3)) return true;
  if (delta <= t_stop_f && a_f > a_r && v_f_star < v_r && t_stop_r < t_stop_f_star) {
    float rel = v_f_star - v_r;
    float sd2 = bm_div(rel * rel, 2.0f * (a_f - a_r)) - front_react + v_r * delta;
    return dist > sd2;
  }
  return dist > sd1;
}

// mcts_state_base.hpp:107-118
__device__ __forceinline__ float reward_from_flags(const bmcts_config& cfg, uint32_t f, float dt) {
  int coll = (f & BMCTS_F_COLLISION_OTHER) != 0;
  int stat = (f & BMCTS_F_STATIC_SAFE_DIST) != 0;
  int dyn = (f & BMCTS_F_DYNAMIC_SAFE_DIST) != 0;
  int driv = (f & BMCTS_F_COLLISION_DRIVABLE) != 0;
  int goal = (f & BMCTS_F_GOAL_REACHED) != 0;
  int oom = (f & BMCTS_F_OUT_OF_MAP) != 0;
  int sd_pred = cfg.static_safe_dist_is_terminal ? dyn : (dyn || stat);
  float safe_dist_reward = ((float)sd_pred * cfg.safe_dist_violated_reward) * dt;
  int coll_pred = coll || (cfg.static_safe_dist_is_terminal ? stat : 0);
  float collision_reward =
      (float)coll_pred * cfg.collision_reward + (float)(driv || oom) * cfg.out_of_drivable_reward;
  float r = bm_max(collision_reward, cfg.collision_reward);
  r = r + (cfg.add_safe_dist ? safe_dist_reward : 0.0f);
  r = r + (float)goal * cfg.goal_reward;
  r = r + cfg.step_reward;
  return r;
}

// mcts_state_base.hpp:120-142
__device__ __forceinline__ void costs_from_flags(const bmcts_config& cfg, uint32_t f, float dt,
                                                 float& c0, float& c1) {
  int coll = (f & BMCTS_F_COLLISION_OTHER) != 0;
  int stat = (f & BMCTS_F_STATIC_SAFE_DIST) != 0;
  int dyn = (f & BMCTS_F_DYNAMIC_SAFE_DIST) != 0;
  int driv = (f & BMCTS_F_COLLISION_DRIVABLE) != 0;
  int goal = (f & BMCTS_F_GOAL_REACHED) != 0;
  int oom = (f & BMCTS_F_OUT_OF_MAP) != 0;
  int sd_pred = cfg.static_safe_dist_is_terminal ? dyn : (dyn || stat);
  float safe_dist_cost = ((float)sd_pred * cfg.safe_dist_violated_cost) * dt;
  int coll_pred = coll || (cfg.static_safe_dist_is_terminal ? stat : 0) ||
                  (cfg.dynamic_safe_dist_is_terminal ? dyn : 0);
  float total = (float)coll_pred * cfg.collision_cost;
  total = total + (float)(driv || oom) * cfg.out_of_drivable_cost;
  total = total + ((cfg.add_safe_dist && !cfg.split_safe_dist_collision) ? safe_dist_cost : 0.0f);
  total = total + (float)goal * cfg.goal_cost;
  if (cfg.split_safe_dist_collision) {
    c0 = cfg.chance_costs ? bm_min(safe_dist_cost, 1.0f) : safe_dist_cost;
    c1 = bm_min(total, cfg.collision_cost) * dt;
  } else {
    c0 = cfg.chance_costs ? bm_min(total, 1.0f) : total;
    c1 = 0.0f;
  }
}

// mcts_state_base.hpp:273-277
__device__ __forceinline__ uint32_t terminal_from_flags(const bmcts_config& cfg, uint32_t f) {
  bool t = (f & (BMCTS_F_COLLISION_OTHER | BMCTS_F_OUT_OF_MAP | BMCTS_F_GOAL_REACHED)) != 0;
  t = t || ((f & BMCTS_F_COLLISION_DRIVABLE) && cfg.out_of_drivable_is_terminal);
  t = t || ((f & BMCTS_F_STATIC_SAFE_DIST) && cfg.static_safe_dist_is_terminal);
  t = t || ((f & BMCTS_F_DYNAMIC_SAFE_DIST) && cfg.dynamic_safe_dist_is_terminal);
  return t ? (uint32_t)BMCTS_F_TERMINAL : 0u;
}

// Euler steps of one macro action (BehaviorMotionPrimitives::IntegrationTimeDelta; oracle
// oracle_ego_substeps): ceil(Dt / delta) with 1e-3 of slack, or num_substeps when delta <= 0
__device__ __forceinline__ int ego_substeps(const bmcts_config& cfg, float Dt) {
  if (!(cfg.ego_integration_dt > 0.0f)) return cfg.num_substeps;
  float k = ceilf(bm_div(Dt, cfg.ego_integration_dt) - 1.0e-3f);
  k = bm_clamp(k, 1.0f, 1024.0f);
  return (int)k;
}

struct IdmParams {
  float T, s0, a_max, v0, b;
};

__device__ __forceinline__ IdmParams sample_idm_params(const bmcts_hypothesis& h, uint4 r0,
                                                       uint4 r1) {
  IdmParams p;
  p.T = bm_fma(bm_u01(r0.x), h.hi[0] - h.lo[0], h.lo[0]);
  p.s0 = bm_fma(bm_u01(r0.y), h.hi[1] - h.lo[1], h.lo[1]);
  p.a_max = bm_fma(bm_u01(r0.z), h.hi[2] - h.lo[2], h.lo[2]);
  p.v0 = bm_fma(bm_u01(r0.w), h.hi[3] - h.lo[3], h.lo[3]);
  p.b = bm_fma(bm_u01(r1.x), h.hi[4] - h.lo[4], h.lo[4]);
  return p;  // coolness (r1.y) is validated to be 0: no blend
}

struct IdmLimits {  // per-hypothesis constants hoisted out of the sub-step loop
  float acc_lo, acc_hi, v_min, v_max;
  int exponent;
};

__device__ __forceinline__ IdmLimits idm_limits(const bmcts_hypothesis& h) {
  IdmLimits l;
  l.acc_lo = h.acc_lower_bound;
  l.acc_hi = h.acc_upper_bound;
  l.v_min = h.min_velocity;
  l.v_max = h.max_velocity;
  l.exponent = h.exponent;
  return l;
}

__device__ __forceinline__ float idm_acc(const IdmLimits& h, const IdmParams& p, float inv_v0,
                                         float inv_2sab, bool has_leader, float v, float v_l,
                                         float gap) {
  float free_term = 1.0f - bm_powi(v * inv_v0, h.exponent);
  float acc;
  if (has_leader) {
    float dv = v - v_l;
    float s_star = p.s0 + bm_max(0.0f, bm_fma(v, p.T, (v * dv) * inv_2sab));
    float q = bm_div(s_star, bm_max(gap, kGapEps));
    acc = p.a_max * (free_term - q * q);
  } else {
    acc = p.a_max * free_term;
  }
  return bm_clamp(acc, h.acc_lo, h.acc_hi);
}

// ---------------------------------------------------------------------------
// K4: BehaviorHypothesisIDM::GetProbability (hypothesis/idm/hypothesis_idm.cpp:40-126)
// one block per (agent a, hypothesis h, world w); threads stride over the samples and
// count those that fall into the bucket of the observed action.
constexpr int kLikelihoodHypsPerBlock = 8;  // at most; LikelihoodArgs::hyps_per_block is what a launch uses

struct LikelihoodArgs {
  const float4* pose;     // [A][n_worlds]
  const uint32_t* alive;  // [n_worlds]
  const float* action;    // [A][n_worlds]
  float* prob;            // [(a*H + h)*n_worlds + w]
  const bmcts_scenario* scn;
  int n_worlds, A, C, H;
  int n_samples, n_buckets;
  float lower, upper;
  uint32_t key0, key1;
  int hyps_per_block;  // 1 .. kLikelihoodHypsPerBlock (fewer when the grid would not fill the GPU)
};

// bucket of an acceleration exactly as the reference computes it
// (hypothesis_idm.cpp:92-95,117-120: floor((a - lower) / bucket_size), clamped to the histogram)
__device__ __forceinline__ int likelihood_bucket(float acc, float lower, float bucket_size, int nb) {
  int b = (int)floorf(bm_div(acc - lower, bucket_size));
  if (b > nb - 1) b = nb - 1;
  if (b < 0) b = 0;
  return b;
}

// floats as integers in their numeric order (and back): for a monotone function of a float the
// set of arguments that map to one value is an interval of consecutive such integers
__device__ __forceinline__ int ordered_of(float x) {
  const int i = __float_as_int(x);
  return i >= 0 ? i : (int)(0x80000000u - (unsigned)i);
}
__device__ __forceinline__ float float_of(int o) {
  return __int_as_float(o >= 0 ? o : (int)(0x80000000u - (unsigned)o));
}

// [lo, hi]: the accelerations inside [lower, upper] whose bucket is `target` -- exactly the
// samples the histogram of the reference counts in that bucket.  likelihood_bucket is monotone in
// acc (subtraction, division by a positive size, floor and the clamps all are), so two binary
// searches over the ordered floats find the interval; the per-sample test is then two
// comparisons instead of a division.  Empty interval: lo > hi.
__device__ void likelihood_interval(float lower, float upper, float bucket_size, int nb, int target,
                                    float* lo_out, float* hi_out) {
  const int o_lo = ordered_of(lower), o_hi = ordered_of(upper);
  // The boundaries sit within a few ulps of lower + k * bucket_size: bracket each of them with a
  // window of +-64 ordered floats around that estimate, check the bracket, and search inside it
  // (6 probes instead of 32); a bracket that does not hold falls back to the whole range.
  auto search = [&](int want_above, int k) {  // first x in [lower, upper] with bucket(x) > want_above
    if (want_above >= nb - 1) return o_hi + 1;  // the last bucket has no upper neighbour
    int a = o_lo, b = o_hi + 1;               // answer in [a, b]
    const int est = ordered_of(bm_fma((float)k, bucket_size, lower));
    const int wa = est - 64 > o_lo ? est - 64 : o_lo, wb = est + 64 < o_hi ? est + 64 : o_hi;
    if (wa <= wb) {
      const bool below_ok = wa == o_lo || likelihood_bucket(float_of(wa - 1), lower, bucket_size, nb) <= want_above;
      const bool above_ok = likelihood_bucket(float_of(wb), lower, bucket_size, nb) > want_above;
      if (below_ok && above_ok) {
        a = wa;
        b = wb;
      }
    }
    while (a < b) {
      const int m = a + (int)(((long long)b - a) / 2);
      if (likelihood_bucket(float_of(m), lower, bucket_size, nb) > want_above) b = m; else a = m + 1;
    }
    return a;
  };
  const int first = search(target - 1, target);  // first x with bucket(x) >= target
  int a = search(target, target + 1);            // first x with bucket(x) > target
  const int last = a - 1;
  if (first > o_hi || last < first) {
    *lo_out = 1.0f;
    *hi_out = 0.0f;
  } else {
    *lo_out = float_of(first);
    *hi_out = float_of(last);
  }
}

__global__ void __launch_bounds__(256) likelihood_kernel(const LikelihoodArgs p) {
  __shared__ int s_cur;
  __shared__ float s_gap, s_vl, s_vown;
  __shared__ int s_has_leader, s_skip;
  __shared__ float s_acc_lo, s_acc_hi;  // accelerations that fall into the observed action's bucket
  __shared__ unsigned s_count[kLikelihoodHypsPerBlock];
  const bmcts_scenario& scn = *p.scn;  // read through L1/L2 (few KB, hot)
  const int w = blockIdx.x;
  // a block serves kLikelihoodHypsPerBlock hypotheses of one (world, agent): corridor, leader
  // and the bucket interval do not depend on the hypothesis and are found once
  const int h0 = blockIdx.y * p.hyps_per_block;
  const int h1 = h0 + p.hyps_per_block < p.H ? h0 + p.hyps_per_block : p.H;
  const int a = blockIdx.z;
  const int t = threadIdx.x;
  const uint32_t am = p.alive[w];
  // Prologue on warp 0 (one lane per corridor, then one lane per candidate leader, reduced with
  // shuffles); the other warps wait at two barriers instead of behind a serial thread.
  if (t < 32) {
    // current corridor of agent a (publish(): argmin |d| over in-range corridors, lowest index
    // on ties)
    const float4 q = p.pose[(size_t)a * p.n_worlds + w];
    float m = kBig;
    if (t < p.C) {
      const Proj pr = project(scn.corridors[t], q.x, q.y);
      m = pr.in ? bm_abs(pr.d) : kBig;
    }
    int bc = t < p.C ? t : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, o);
      const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
      const bool take = om < m || (om == m && oc < bc);
      m = take ? om : m;
      bc = take ? oc : bc;
    }
    if (t < kLikelihoodHypsPerBlock) s_count[t] = 0;
    if (t == 0) {
      s_cur = m < kBig ? bc : 0;  // no corridor in range: corridor 0, as the serial scan
      s_vown = q.w;
      int skip = !((am >> a) & 1u);
      const float act = p.action[(size_t)a * p.n_worlds + w];
      if (act > p.upper || act < p.lower) skip = 1;
      s_skip = skip;
      const float bucket_size = bm_div(p.upper - p.lower, (float)p.n_buckets);
      float lo = 1.0f, hi = 0.0f;
      if (!skip) {
        const int target = likelihood_bucket(act, p.lower, bucket_size, p.n_buckets);
        likelihood_interval(p.lower, p.upper, bucket_size, p.n_buckets, target, &lo, &hi);
      }
      s_acc_lo = lo;
      s_acc_hi = hi;
    }
  }
  __syncthreads();
  if (s_skip) {
    for (int h = h0 + t; h < h1; h += blockDim.x)
      p.prob[((size_t)a * p.H + h) * p.n_worlds + w] = 0.0f;
    return;
  }
  const int cur = s_cur;
  if (t < 32) {
    // BaseIDM::CalcRelativeValues (find_leader): nearest alive agent ahead in the corridor,
    // lowest slot on equal distances
    const float hwc = scn.corridors[cur].half_width;
    const float4 qa = p.pose[(size_t)a * p.n_worlds + w];
    const float s_own = project(scn.corridors[cur], qa.x, qa.y).s;
    float ds = kBig, vj = 0.0f;
    int j = 0x7fffffff;
    if (t < p.A && t != a && ((am >> t) & 1u)) {
      const float4 qj = p.pose[(size_t)t * p.n_worlds + w];
      const Proj pr = project(scn.corridors[cur], qj.x, qj.y);
      const float dj = pr.in ? pr.d : kBig;
      const float d = pr.s - s_own;
      if (bm_abs(dj) <= hwc && d > 0.0f) {
        ds = d;
        j = t;
        vj = qj.w;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ods = __shfl_xor_sync(0xffffffffu, ds, o);
      const int oj = __shfl_xor_sync(0xffffffffu, j, o);
      const float ov = __shfl_xor_sync(0xffffffffu, vj, o);
      const bool take = ods < ds || (ods == ds && oj < j);
      ds = take ? ods : ds;
      j = take ? oj : j;
      vj = take ? ov : vj;
    }
    if (t == 0) {
      const bool found = ds < kBig;
      s_has_leader = found;
      s_gap = found ? ds - ((scn.agent_length[a] - scn.agent_rear[a]) + scn.agent_rear[j]) : 0.0f;
      s_vl = found ? vj : 0.0f;
    }
  }
  __syncthreads();
  const bool has_leader = s_has_leader != 0;
  const float gap = s_gap, v_l = s_vl, v_own = s_vown;
  const float acc_lo = s_acc_lo, acc_hi = s_acc_hi;
#pragma unroll 1
  for (int h = h0; h < h1; ++h) {
    const bmcts_hypothesis hy = scn.hypotheses[h];
    const IdmLimits hl = idm_limits(hy);
    // Exact skips, as in the rollout kernels: a Philox block none of whose words reaches a ranged
    // parameter is not drawn (u * 0 + lo = lo whatever u), and the reciprocals of parameters that
    // are FixedValues are the same for every sample.
    const bool need0 = hy.hi[0] != hy.lo[0] || hy.hi[1] != hy.lo[1] || hy.hi[2] != hy.lo[2] ||
                       hy.hi[3] != hy.lo[3];
    const bool need1 = hy.hi[4] != hy.lo[4];
    const bool fixed_v0 = hy.hi[3] == hy.lo[3];
    const bool fixed_ab = hy.hi[2] == hy.lo[2] && hy.hi[4] == hy.lo[4];
    const float inv_v0_fixed = bm_div(1.0f, bm_max(hy.lo[3], 1.0e-3f));
    const float inv_2sab_fixed = bm_div(1.0f, 2.0f * bm_sqrt(bm_max(hy.lo[2] * hy.lo[4], 1.0e-6f)));
    unsigned cnt = 0;
#pragma unroll 1
    for (int smp = t; smp < p.n_samples; smp += blockDim.x) {
      uint64_t id = (uint64_t)(uint32_t)smp | ((uint64_t)(uint32_t)w << 32);
      uint4 r0 = make_uint4(0u, 0u, 0u, 0u), r1 = r0;
      if (need0) r0 = draw(p.key0, p.key1, id, (uint32_t)a, BMCTS_RNG_LIKELIHOOD, 2u * (uint32_t)h);
      if (need1)
        r1 = draw(p.key0, p.key1, id, (uint32_t)a, BMCTS_RNG_LIKELIHOOD, 2u * (uint32_t)h + 1u);
      IdmParams ip = sample_idm_params(hy, r0, r1);
      const float inv_v0 = fixed_v0 ? inv_v0_fixed : bm_div(1.0f, bm_max(ip.v0, 1.0e-3f));
      const float inv_2sab =
          fixed_ab ? inv_2sab_fixed
                   : bm_div(1.0f, 2.0f * bm_sqrt(bm_max(ip.a_max * ip.b, 1.0e-6f)));
      const float acc = idm_acc(hl, ip, inv_v0, inv_2sab, has_leader, v_own, v_l, gap);
      // the observed action's bucket, as an interval of accelerations (likelihood_interval)
      cnt += (acc >= acc_lo && acc <= acc_hi);
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((t & 31) == 0 && cnt) atomicAdd(&s_count[h - h0], cnt);
  }
  __syncthreads();
  for (int h = h0 + t; h < h1; h += blockDim.x)
    p.prob[((size_t)a * p.H + h) * p.n_worlds + w] = bm_div((float)s_count[h - h0], (float)p.n_samples);
}

// elementary-function probe (bmcts_math_probe): out[i] = f(a[i], b[i])
__global__ void math_probe_kernel(int op, const float* a, const float* b, int n, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = a[i], y = b[i];
  float s, c, r = 0.0f;
  switch (op) {
    case BMCTS_MATH_SIN: bm_sincos(x, &s, &c); r = s; break;
    case BMCTS_MATH_COS: bm_sincos(x, &s, &c); r = c; break;
    case BMCTS_MATH_ATAN2: r = bm_atan2(x, y); break;
    case BMCTS_MATH_TAN_SMALL: r = bm_tan_small(x); break;
    case BMCTS_MATH_DIV: r = bm_div(x, y); break;
    case BMCTS_MATH_SQRT: r = bm_sqrt(x); break;
    case BMCTS_MATH_MIN: r = bm_min(x, y); break;
    case BMCTS_MATH_MAX: r = bm_max(x, y); break;
    case BMCTS_MATH_NORM_ANGLE: r = bm_norm_angle(x); break;
    case BMCTS_MATH_FMA_AAB: r = bm_fma(x, x, y); break;
    case BMCTS_MATH_POW4: r = bm_powi(x, 4); break;
    default: break;
  }
  out[i] = r;
}

}  // namespace bmcts
