// rollout_kernel.cuh -- sm_100a kernels of the MCTS leaf-evaluation path.
//
// One kernel implements the three batched entry points of the C ABI:
//   step    (tree expansion)      = MctsState*::execute + plan_action_current_hypothesis
//                                   mcts_state/mcts_state_hypothesis.cpp:62-114,
//                                   mcts_state/mcts_state_cooperative.cpp:45-116
//   rollout (leaf evaluation)     = mamcts RandomHeuristic::calculate_heuristic_values
//                                    looping the same step
//   expand + rollout              = both, fused
// (paths relative to /root/reference/bark_mcts/models/behavior/).
//
// Mapping ("agent-per-warp"): a block owns 32 rollouts and has A warps.  Warp a
// is agent slot a, lane r is rollout r of the block.  Consequences:
//   * a warp never diverges between the ego's single-track macro action and the
//     others' IDM hypothesis model -- they live in different warps;
//   * all 32 lanes are busy whatever A is (no 20-of-32 packing loss);
//   * batch arrays are agent-major SoA, so warp a reads 32 consecutive float4
//     poses (512 B, fully coalesced);
//   * agents of one rollout exchange their published state through shared memory
//     laid out [field][agent][lane] -> bank = lane, conflict free; lane geometry,
//     road polygon, goal, shapes and the hypothesis table are staged in shared
//     memory once per block.
// Every float operation mirrors oracle/bmcts_oracle.c operation by operation
// (compile with --fmad=false; FMAs only where bm_fma is written).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "bmcts_c.h"
#include "bmcts_math.h"

namespace bmcts {

constexpr int kLanes = 32;
constexpr float kBig = 1.0e30f;
constexpr float kGapEps = 1.0e-3f;
constexpr uint32_t kGoalFailBit = 1u << 16;  // scratch bit: a goal-polygon corner is outside

// published per-agent fields in shared memory
enum { F_X = 0, F_Y, F_C, F_S, F_V, F_TH, F_PROJ };  // F_PROJ + 2*c = s, +1 = d_eff

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
};

__host__ __device__ inline size_t smem_align16(size_t x) { return (x + 15) & ~size_t(15); }

__host__ inline size_t rollout_smem_bytes(int A, int C) {
  size_t b = smem_align16(sizeof(bmcts_scenario)) + smem_align16(sizeof(bmcts_config));
  b += sizeof(float) * (size_t)(F_PROJ + 2 * C) * A * kLanes;  // pub
  b += sizeof(float) * (size_t)A * kLanes;                     // rw (own rewards, cooperative)
  b += sizeof(uint32_t) * kLanes * 3;                          // amask, flags, ctrl
  return b;
}

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
#pragma unroll 1
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
  float t_stop_f_star = bm_div(-v_f_star, a_r);
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
template <bool COOP>
__global__ void __launch_bounds__(1024) rollout_kernel(const KernelArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int A = p.A, C = p.C;
  bmcts_scenario* scn_w = reinterpret_cast<bmcts_scenario*>(smem_raw);
  bmcts_config* cfg_w =
      reinterpret_cast<bmcts_config*>(smem_raw + smem_align16(sizeof(bmcts_scenario)));
  float* pub = reinterpret_cast<float*>(smem_raw + smem_align16(sizeof(bmcts_scenario)) +
                                        smem_align16(sizeof(bmcts_config)));
  float* rw = pub + (size_t)(F_PROJ + 2 * C) * A * kLanes;
  uint32_t* amask = reinterpret_cast<uint32_t*>(rw + (size_t)A * kLanes);
  uint32_t* flags = amask + kLanes;
  uint32_t* ctrl = flags + kLanes;

  // stage scenario + config in shared memory (16-byte vector copies)
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.scn);
    uint4* dst = reinterpret_cast<uint4*>(scn_w);
    for (int i = threadIdx.x; i < (int)(sizeof(bmcts_scenario) / 16); i += blockDim.x)
      dst[i] = src[i];
    const uint32_t* s2 = reinterpret_cast<const uint32_t*>(p.cfg);
    uint32_t* d2 = reinterpret_cast<uint32_t*>(cfg_w);
    for (int i = threadIdx.x; i < (int)(sizeof(bmcts_config) / 4); i += blockDim.x) d2[i] = s2[i];
  }
  const bmcts_scenario& scn = *scn_w;
  const bmcts_config& cfg = *cfg_w;

  const int k = threadIdx.x >> 5;  // agent slot
  const int r = threadIdx.x & 31;  // rollout lane
  const int rid = blockIdx.x * kLanes + r;
  const bool valid = rid < p.n;
  const size_t gi = (size_t)k * p.n + (valid ? rid : 0);  // agent-major element index
  const bool macro = COOP || k == 0;

#define PUB(f, a) pub[((f) * A + (a)) * kLanes + r]

  float x = 0.0f, y = 0.0f, th = 0.0f, v = 0.0f;
  bool alive = false;
  int hyp = 0;
  int depth = 1;
  if (valid) {
    float4 q = p.pose[gi];
    x = q.x;
    y = q.y;
    th = q.z;
    v = q.w;
    uint32_t am = p.alive[rid];
    alive = (am >> k) & 1u;
    hyp = p.hyp ? (int)p.hyp[gi] : 0;
    depth = p.depth ? (int)p.depth[rid] : 1;
    if (k == 0) amask[r] = am;
  } else if (k == 0) {
    amask[r] = 0;
  }
  if (k == 0) {
    flags[r] = 0;
    ctrl[r] = valid ? 0u : 1u;
  }
  __syncthreads();

  // own published quantities kept in registers
  float cth = 1.0f, sth = 0.0f, s_cur = 0.0f;
  int cur = 0;
  // goal (Frenet limits) ingredients of the own projection
  float goal_d = 0.0f, goal_ang = 0.0f;
  bool goal_in = false;

  auto publish = [&]() {
    if (!alive) return;
    bm_sincos(th, &sth, &cth);
    PUB(F_X, k) = x;
    PUB(F_Y, k) = y;
    PUB(F_C, k) = cth;
    PUB(F_S, k) = sth;
    PUB(F_V, k) = v;
    PUB(F_TH, k) = th;
    float best = kBig;
    int bc = 0;
    for (int c = 0; c < C; ++c) {
      Proj pr = project(scn.corridors[c], x, y);
      float m = pr.in ? bm_abs(pr.d) : kBig;
      PUB(F_PROJ + 2 * c, k) = pr.s;
      PUB(F_PROJ + 2 * c + 1, k) = pr.in ? pr.d : kBig;
      if (m < best) {
        best = m;
        bc = c;
        s_cur = pr.s;
      }
      if (scn.goal.kind == BMCTS_GOAL_FRENET_LIMITS && c == scn.goal.corridor) {
        goal_d = pr.d;
        goal_in = pr.in;
        goal_ang = bm_norm_angle(th - scn.corridors[c].angle[pr.seg]);
      }
    }
    if (best == kBig) s_cur = PUB(F_PROJ, k);  // no corridor in range: corridor 0 (oracle: cur = 0)
    cur = bc;
  };

  // nearest alive agents ahead of / behind the own agent inside its corridor
  int jf = -1, jr = -1;
  float ds_f = kBig, ds_r = kBig;
  const bool need_rear = cfg.add_safe_dist && cfg.dyn_to_rear;
  auto neighbours = [&](uint32_t am) {
    jf = -1;
    jr = -1;
    ds_f = kBig;
    ds_r = kBig;
    if (!alive) return;
    const float hwc = scn.corridors[cur].half_width;
    for (int j = 0; j < A; ++j) {
      if (j == k || !((am >> j) & 1u)) continue;
      float dj = PUB(F_PROJ + 2 * cur + 1, j);
      if (!(bm_abs(dj) <= hwc)) continue;
      float ds = PUB(F_PROJ + 2 * cur, j) - s_cur;
      if (ds > 0.0f && ds < ds_f) {
        ds_f = ds;
        jf = j;
      }
      float dr = -ds;
      if (need_rear && dr > 0.0f && dr < ds_r) {
        ds_r = dr;
        jr = j;
      }
    }
  };

  publish();
  __syncthreads();
  neighbours(amask[r]);

  bool active = valid;
  bool expand = p.do_expand != 0;
  int n = 0;  // rollout steps executed
  float disc = cfg.discount_factor;
  float R = 0.0f, C0 = 0.0f, C1 = 0.0f, L = 0.0f, Ra = 0.0f;
  uint32_t f_any = 0, f_last = 0;
  int agent_steps = 0;
  const uint64_t rollout_id = p.first_rollout_id + (uint64_t)(valid ? rid : 0);

  if (!expand && (!p.do_rollout || cfg.max_rollout_steps <= 0)) active = false;

  while (__syncthreads_or(active)) {
    // ------------------------------------------------------------ predict()
    const uint32_t pre_mask = amask[r];
    const float Dt = p.dt_table ? p.dt_table[depth < BMCTS_MAX_DEPTH ? depth : BMCTS_MAX_DEPTH - 1]
                                : cfg.prediction_k;
    float nx = x, ny = y, nth = th, nv = v, la = 0.0f;
    bool nalive = alive;
    if (active && alive) {
      const int nsub = cfg.num_substeps;
      const float dt = bm_div(Dt, (float)nsub);
      if (macro) {
        // BehaviorMPMacroActions primitive on the single-track model (spec item 8)
        int action;
        if (expand) {
          action = p.action[gi];
        } else {
          uint4 rr = draw(p.key0, p.key1, rollout_id, (uint32_t)k, BMCTS_RNG_ROLLOUT_ACTION,
                          (uint32_t)n);
          action = (int)bm_bounded(rr.x, (uint32_t)scn.n_ego_actions);
        }
        const bmcts_ego_action act = scn.ego_actions[action];
        const float acc = bm_clamp(act.acceleration, cfg.lon_acc_min, cfg.lon_acc_max);
        int tgt = cur;
        if (act.type == BMCTS_ACT_CHANGE_LEFT && scn.corridors[tgt].left >= 0)
          tgt = scn.corridors[tgt].left;
        else if (act.type == BMCTS_ACT_CHANGE_RIGHT && scn.corridors[tgt].right >= 0)
          tgt = scn.corridors[tgt].right;
        const bmcts_corridor& corr = scn.corridors[tgt];
        const float inv_wb = bm_div(1.0f, cfg.wheel_base);
        for (int q = 0; q < nsub; ++q) {
          float sn, cs;
          bm_sincos(nth, &sn, &cs);
          float fx = bm_fma(cfg.wheel_base, cs, nx);
          float fy = bm_fma(cfg.wheel_base, sn, ny);
          Proj pr = project(corr, fx, fy);
          float th_err = bm_norm_angle(corr.angle[pr.seg] - nth);
          float delta = th_err + bm_atan2(-(cfg.crosstrack_gain * pr.d), nv);
          delta = bm_clamp(delta, -cfg.delta_max, cfg.delta_max);
          float tn = bm_tan_small(delta);
          float dv = dt * nv;
          nx = bm_fma(dv, cs, nx);
          ny = bm_fma(dv, sn, ny);
          nth = bm_fma(dv, tn * inv_wb, nth);
          nv = bm_clamp(bm_fma(dt, acc, nv), 0.0f, cfg.ego_max_velocity);
        }
        nth = bm_norm_angle(nth);
        la = acc;
      } else {
        // BehaviorIDMStochastic::Plan under the sampled hypothesis (spec items 5-7)
        const bmcts_hypothesis& h = scn.hypotheses[hyp];
        uint4 r0, r1;
        if (expand) {
          uint64_t id = p.draw_id ? p.draw_id[gi] : 0ull;
          r0 = draw(p.key0, p.key1, id, (uint32_t)k, BMCTS_RNG_TREE_PARAMS, 0u);
          r1 = draw(p.key0, p.key1, id, (uint32_t)k, BMCTS_RNG_TREE_PARAMS, 1u);
        } else {
          r0 = draw(p.key0, p.key1, rollout_id, (uint32_t)k, BMCTS_RNG_ROLLOUT_PARAMS,
                    2u * (uint32_t)n);
          r1 = draw(p.key0, p.key1, rollout_id, (uint32_t)k, BMCTS_RNG_ROLLOUT_PARAMS,
                    2u * (uint32_t)n + 1u);
        }
        const IdmParams ip = sample_idm_params(h, r0, r1);
        const IdmLimits hl = idm_limits(h);
        const bool has_leader = jf >= 0;
        float gap = 0.0f, v_l = 0.0f;
        if (has_leader) {
          gap = ds_f - ((scn.agent_length[k] - scn.agent_rear[k]) + scn.agent_rear[jf]);
          v_l = PUB(F_V, jf);
        }
        const float inv_v0 = bm_div(1.0f, bm_max(ip.v0, 1.0e-3f));
        const float inv_2sab = bm_div(1.0f, 2.0f * bm_sqrt(bm_max(ip.a_max * ip.b, 1.0e-6f)));
        const float half_dt2 = (0.5f * dt) * dt;
        float trav = 0.0f, acc = 0.0f;
        for (int q = 0; q < nsub; ++q) {
          acc = idm_acc(hl, ip, inv_v0, inv_2sab, has_leader, nv, v_l, gap);
          float ds = bm_max(0.0f, bm_fma(acc, half_dt2, nv * dt));
          nv = bm_clamp(bm_fma(acc, dt, nv), hl.v_min, hl.v_max);
          gap = gap + bm_fma(v_l, dt, -ds);
          trav = trav + ds;
        }
        const bmcts_corridor& corr = scn.corridors[cur];
        float s_new = bm_min(s_cur + trav, corr.length);
        const int nseg = corr.n_pts - 1;
        int sg = 0;
        for (int q = 1; q < nseg; ++q)
          if (corr.s0[q] <= s_new) sg = q;
        float ls = s_new - corr.s0[sg];
        nx = bm_fma(ls, corr.tx[sg], corr.px[sg]);
        ny = bm_fma(ls, corr.ty[sg], corr.py[sg]);
        nth = corr.angle[sg];
        la = acc;
      }
      // World::RemoveInvalidAgents (spec item 9)
      nalive = point_in_polygon(scn.road_x, scn.road_y, scn.n_road_pts, nx, ny);
    }
    if (k == 0 && active && !expand) agent_steps += __popc(pre_mask);
    __syncthreads();  // every read of the pre-step snapshot is done
    if (active && alive) {
      x = nx;
      y = ny;
      th = nth;
      v = nv;
      if (!nalive) {
        alive = false;
        atomicAnd(&amask[r], ~(1u << k));
      }
    }
    if (active) publish();
    __syncthreads();  // post-step world published
    const uint32_t post_mask = amask[r];
    neighbours(post_mask);

    // --------------------------------------- mcts_observed_world_evaluation
    uint32_t f_own = 0;  // COOP: own perspective; else: ego warp's private part
    if (active) {
      if (COOP) {
        if (!alive) {
          f_own |= BMCTS_F_OUT_OF_MAP;
        } else {
          Box bd = agent_box(scn, k, x, y, cth, sth, cfg.drivable_lat_safe_dist,
                             cfg.drivable_lon_safe_dist);
          bool within = true;
          for (int q = 0; q < 4; ++q)
            within = within && corner_in_polygon(bd, q, scn.road_x, scn.road_y, scn.n_road_pts);
          if (!within) f_own |= BMCTS_F_COLLISION_DRIVABLE;
          Box be = agent_box(scn, k, x, y, cth, sth, 0.0f, 0.0f);
          Box bs = agent_box(scn, k, x, y, cth, sth, cfg.static_lat_safe_dist,
                             cfg.static_lon_safe_dist);
          for (int j = 0; j < A; ++j) {
            if (j == k || !((post_mask >> j) & 1u)) continue;
            Box bj = agent_box(scn, j, PUB(F_X, j), PUB(F_Y, j), PUB(F_C, j), PUB(F_S, j), 0.0f,
                               0.0f);
            if (boxes_overlap(be, bj)) f_own |= BMCTS_F_COLLISION_OTHER;
            if (cfg.add_safe_dist && boxes_overlap(bs, bj)) f_own |= BMCTS_F_STATIC_SAFE_DIST;
          }
          if (scn.goal.kind == BMCTS_GOAL_POLYGON) {
            bool in_goal = true;
            for (int q = 0; q < 4; ++q)
              in_goal = in_goal && corner_in_polygon(be, q, scn.goal.px, scn.goal.py, scn.goal.n_pts);
            if (in_goal) f_own |= BMCTS_F_GOAL_REACHED;
          }
        }
      } else {
        const bool ego_alive = post_mask & 1u;
        if (ego_alive) {
          // the ego's 4 + 4 corner-in-polygon tests are spread over the other warps
          const float ex = PUB(F_X, 0), ey = PUB(F_Y, 0), ec = PUB(F_C, 0), es = PUB(F_S, 0);
          const int n_items = (scn.goal.kind == BMCTS_GOAL_POLYGON) ? 8 : 4;
          // item t belongs to warp 1 + t % (A-1) (warp 0 when the ego is alone)
          const int t_first = (A > 1) ? k - 1 : 0;
          const int t_stride = (A > 1) ? A - 1 : 1;
          for (int t = t_first; t >= 0 && t < n_items; t += t_stride) {
            if (t < 4) {
              Box bd = agent_box(scn, 0, ex, ey, ec, es, cfg.drivable_lat_safe_dist,
                                 cfg.drivable_lon_safe_dist);
              if (!corner_in_polygon(bd, t, scn.road_x, scn.road_y, scn.n_road_pts))
                atomicOr(&flags[r], (uint32_t)BMCTS_F_COLLISION_DRIVABLE);
            } else {
              Box be = agent_box(scn, 0, ex, ey, ec, es, 0.0f, 0.0f);
              if (!corner_in_polygon(be, t - 4, scn.goal.px, scn.goal.py, scn.goal.n_pts))
                atomicOr(&flags[r], kGoalFailBit);
            }
          }
          if (k > 0 && alive) {
            // EvaluatorCollisionEgoAgent / EvaluatorStaticSafeDist pair (ego, k)
            Box bj = agent_box(scn, k, x, y, cth, sth, 0.0f, 0.0f);
            Box be = agent_box(scn, 0, ex, ey, ec, es, 0.0f, 0.0f);
            uint32_t hit = 0;
            if (boxes_overlap(be, bj)) hit |= BMCTS_F_COLLISION_OTHER;
            if (cfg.add_safe_dist) {
              Box bs = agent_box(scn, 0, ex, ey, ec, es, cfg.static_lat_safe_dist,
                                 cfg.static_lon_safe_dist);
              if (boxes_overlap(bs, bj)) hit |= BMCTS_F_STATIC_SAFE_DIST;
            }
            if (hit) atomicOr(&flags[r], hit);
          }
        }
      }
      // parts every perspective agent evaluates for itself (ego warp unless COOP)
      if ((COOP || k == 0) && alive) {
        if (scn.goal.kind == BMCTS_GOAL_FRENET_LIMITS) {
          const bmcts_goal& g = scn.goal;
          bool ok = goal_in && goal_d <= g.max_lat_left && -goal_d <= g.max_lat_right &&
                    goal_ang <= g.max_angle_left && -goal_ang <= g.max_angle_right &&
                    v >= g.vel_lo && v <= g.vel_hi;
          if (ok) f_own |= BMCTS_F_GOAL_REACHED;
        }
        if (cfg.add_safe_dist) {
          bool safe = true;
          if (jf >= 0) {
            float dist = ds_f - ((scn.agent_length[k] - scn.agent_rear[k]) + scn.agent_rear[jf]);
            safe = safe && safe_dist_longitudinal(PUB(F_V, jf), v, dist, -cfg.dyn_max_ego_decel,
                                                  -cfg.dyn_max_other_decel,
                                                  cfg.dyn_reaction_time_ego);
          }
          if (jr >= 0 && cfg.dyn_to_rear) {
            float dist = ds_r - ((scn.agent_length[jr] - scn.agent_rear[jr]) + scn.agent_rear[k]);
            safe = safe && safe_dist_longitudinal(v, PUB(F_V, jr), dist, -cfg.dyn_max_other_decel,
                                                  -cfg.dyn_max_ego_decel,
                                                  cfg.dyn_reaction_time_others);
          }
          if (!safe) f_own |= BMCTS_F_DYNAMIC_SAFE_DIST;
        }
      }
      if (!COOP && k == 0 && !alive) f_own |= BMCTS_F_OUT_OF_MAP;
      if (COOP) {
        f_own |= terminal_from_flags(cfg, f_own);
        // own reward; only agents still in the map are counted (cooperative.cpp:76-84)
        rw[k * kLanes + r] = (alive || k == 0) ? reward_from_flags(cfg, f_own, Dt) : 0.0f;
      }
    }
    __syncthreads();  // flags / own rewards complete

    // ------------------------------------------------- generate_next_state
    float reward_k = 0.0f, c0 = 0.0f, c1 = 0.0f;
    uint32_t f_ego = 0;
    if (active) {
      if (COOP) {
        // mcts_state_cooperative.cpp:86-111
        const float r_ego = rw[r];
        float sum = r_ego;
        for (int j = 1; j < A; ++j)
          if ((post_mask >> j) & 1u) sum = sum + rw[j * kLanes + r];
        const uint32_t others = ~1u;
        const int n_in = __popc(post_mask & others);
        const uint32_t missing = pre_mask & ~post_mask & others;
        const float cf = cfg.cooperation_factor;
        const bool was_alive = (pre_mask >> k) & 1u;
        if (k == 0) {
          int n_map = n_in + __popc(missing);
          float rest = sum - r_ego;
          float mix = (1.0f - cf) * r_ego + cf * rest;
          reward_k = bm_div(1.0f, (float)(n_map + 1)) * mix;
          c0 = -reward_k;
          f_ego = f_own;
        } else if (was_alive) {
          int n_map = n_in + __popc(missing & ((2u << k) - 1u));
          float own = ((post_mask >> k) & 1u) ? rw[k * kLanes + r] : 0.0f;
          float rest = sum - own;
          float mix = (1.0f - cf) * own + cf * rest;
          reward_k = bm_div(1.0f, (float)(n_map + 1)) * mix;
        }
      } else if (k == 0) {
        uint32_t f = (flags[r] | f_own);
        if (alive && scn.goal.kind == BMCTS_GOAL_POLYGON && !(f & kGoalFailBit))
          f |= BMCTS_F_GOAL_REACHED;
        f &= 0xffu;
        f |= terminal_from_flags(cfg, f);
        flags[r] = 0;
        f_ego = f;
        reward_k = reward_from_flags(cfg, f, Dt);
        costs_from_flags(cfg, f, Dt, c0, c1);
      }
    }
    bool finished = false;
    if (active) {
      if (expand) {
        // step outputs
        if (p.next_pose) p.next_pose[gi] = make_float4(x, y, th, v);
        if (p.reward) p.reward[gi] = reward_k;
        if (p.last_action) p.last_action[gi] = la;
        if (k == 0) {
          if (p.next_alive) p.next_alive[rid] = post_mask;
          if (p.cost) {
            p.cost[rid] = c0;
            p.cost[(size_t)p.n + rid] = c1;
          }
          if (p.flags) p.flags[rid] = f_ego;
          if (!p.do_rollout || (f_ego & BMCTS_F_TERMINAL) || cfg.max_rollout_steps <= 0) {
            // RandomHeuristic on a terminal child: zero estimate, flags of the expansion step
            finished = true;
            f_last = f_ego;
            f_any = f_ego;
          }
        }
      } else {
        Ra = bm_fma(disc, reward_k, Ra);
        if (k == 0) {
          R = bm_fma(disc, reward_k, R);
          C0 = bm_fma(disc, c0, C0);
          C1 = bm_fma(disc, c1, C1);
          L = L + Dt;
          f_last = f_ego;
          f_any |= f_ego;
          if ((f_ego & BMCTS_F_TERMINAL) || n + 1 >= cfg.max_rollout_steps) finished = true;
        }
        disc = disc * cfg.discount_factor;
        n++;
      }
      depth++;
      if (k == 0 && finished) ctrl[r] = 1u;
    }
    expand = false;
    __syncthreads();  // ctrl visible
    if (active && ctrl[r]) {
      active = false;
      if (p.do_rollout) {
        if (k == 0 && p.result) {
          bmcts_rollout_result res;
          res.ret_ego = R;
          res.cost0 = C0;
          res.cost1 = C1;
          res.step_len = L;
          res.flags = (f_last & 0xffu) | ((f_any & 0xffu) << 8);
          res.n_steps = n;
          res.agent_steps = agent_steps;
          res.final_alive = amask[r];
          float4* dst = reinterpret_cast<float4*>(&p.result[rid]);
          dst[0] = make_float4(res.ret_ego, res.cost0, res.cost1, res.step_len);
          dst[1] = make_float4(__uint_as_float(res.flags), __int_as_float(res.n_steps),
                               __int_as_float(res.agent_steps), __uint_as_float(res.final_alive));
        }
        if (p.ret_agents) p.ret_agents[gi] = Ra;
        if (p.final_pose) p.final_pose[gi] = make_float4(x, y, th, v);
      }
    }
  }
#undef PUB
}

// ---------------------------------------------------------------------------
// K4: BehaviorHypothesisIDM::GetProbability (hypothesis/idm/hypothesis_idm.cpp:40-126)
// one block per (agent a, hypothesis h, world w); threads stride over the samples and
// count those that fall into the bucket of the observed action.
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
};

__global__ void __launch_bounds__(256) likelihood_kernel(const LikelihoodArgs p) {
  __shared__ float s_s[BMCTS_MAX_AGENTS], s_d[BMCTS_MAX_AGENTS], s_v[BMCTS_MAX_AGENTS];
  __shared__ int s_cur;
  __shared__ float s_gap, s_vl, s_vown;
  __shared__ int s_has_leader, s_skip, s_bucket;
  __shared__ unsigned s_count;
  const bmcts_scenario& scn = *p.scn;  // read through L1/L2 (few KB, hot)
  const int w = blockIdx.x;
  const int h = blockIdx.y;
  const int a = blockIdx.z;
  const int t = threadIdx.x;
  const uint32_t am = p.alive[w];
  const size_t out_idx = ((size_t)a * p.H + h) * p.n_worlds + w;
  if (t == 0) {
    s_count = 0;
    s_skip = 0;
    // current corridor of agent a (publish(): argmin |d| over in-range corridors)
    float4 q = p.pose[(size_t)a * p.n_worlds + w];
    float best = kBig;
    int bc = 0;
    for (int c = 0; c < p.C; ++c) {
      Proj pr = project(scn.corridors[c], q.x, q.y);
      float m = pr.in ? bm_abs(pr.d) : kBig;
      if (m < best) {
        best = m;
        bc = c;
      }
    }
    s_cur = bc;
    s_vown = q.w;
    if (!((am >> a) & 1u)) s_skip = 1;
    const float act = p.action[(size_t)a * p.n_worlds + w];
    if (act > p.upper || act < p.lower) s_skip = 1;
    const float bucket_size = bm_div(p.upper - p.lower, (float)p.n_buckets);
    int b = (int)floorf(bm_div(act - p.lower, bucket_size));
    if (b > p.n_buckets - 1) b = p.n_buckets - 1;
    if (b < 0) b = 0;
    s_bucket = b;
  }
  __syncthreads();
  if (s_skip) {
    if (t == 0) p.prob[out_idx] = 0.0f;
    return;
  }
  const int cur = s_cur;
  if (t < p.A) {
    float4 q = p.pose[(size_t)t * p.n_worlds + w];
    Proj pr = project(scn.corridors[cur], q.x, q.y);
    s_s[t] = pr.s;
    s_d[t] = pr.in ? pr.d : kBig;
    s_v[t] = q.w;
  }
  __syncthreads();
  if (t == 0) {
    // BaseIDM::CalcRelativeValues (find_leader)
    float best = kBig;
    int lead = -1;
    const float hwc = scn.corridors[cur].half_width;
    for (int j = 0; j < p.A; ++j) {
      if (j == a || !((am >> j) & 1u)) continue;
      if (!(bm_abs(s_d[j]) <= hwc)) continue;
      float ds = s_s[j] - s_s[a];
      if (ds > 0.0f && ds < best) {
        best = ds;
        lead = j;
      }
    }
    s_has_leader = lead >= 0;
    s_gap = 0.0f;
    s_vl = 0.0f;
    if (lead >= 0) {
      s_gap = best - ((scn.agent_length[a] - scn.agent_rear[a]) + scn.agent_rear[lead]);
      s_vl = s_v[lead];
    }
  }
  __syncthreads();
  const bmcts_hypothesis hy = scn.hypotheses[h];
  const IdmLimits hl = idm_limits(hy);
  const bool has_leader = s_has_leader != 0;
  const float gap = s_gap, v_l = s_vl, v_own = s_vown;
  const float bucket_size = bm_div(p.upper - p.lower, (float)p.n_buckets);
  const int target = s_bucket;
  unsigned cnt = 0;
  for (int smp = t; smp < p.n_samples; smp += blockDim.x) {
    uint64_t id = (uint64_t)(uint32_t)smp | ((uint64_t)(uint32_t)w << 32);
    uint4 r0 = draw(p.key0, p.key1, id, (uint32_t)a, BMCTS_RNG_LIKELIHOOD, 2u * (uint32_t)h);
    uint4 r1 = draw(p.key0, p.key1, id, (uint32_t)a, BMCTS_RNG_LIKELIHOOD, 2u * (uint32_t)h + 1u);
    IdmParams ip = sample_idm_params(hy, r0, r1);
    float inv_v0 = bm_div(1.0f, bm_max(ip.v0, 1.0e-3f));
    float inv_2sab = bm_div(1.0f, 2.0f * bm_sqrt(bm_max(ip.a_max * ip.b, 1.0e-6f)));
    float acc = idm_acc(hl, ip, inv_v0, inv_2sab, has_leader, v_own, v_l, gap);
    if (acc < p.lower || acc > p.upper) continue;
    int b = (int)floorf(bm_div(acc - p.lower, bucket_size));
    if (b > p.n_buckets - 1) b = p.n_buckets - 1;
    if (b < 0) b = 0;
    cnt += (b == target);
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((t & 31) == 0 && cnt) atomicAdd(&s_count, cnt);
  __syncthreads();
  if (t == 0) p.prob[out_idx] = bm_div((float)s_count, (float)p.n_samples);
}

}  // namespace bmcts
