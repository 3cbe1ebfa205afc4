// rollout_lane_kernel.cuh -- sm_100a rollout kernel of the hypothesis / risk-constraint
// states, "rollout-per-lane-group" mapping.
//
// Replaces, for MctsStateHypothesis<void> and MctsStateRiskConstraint,
//   step    = MctsStateHypothesis<T>::execute + plan_action_current_hypothesis
//             (mcts_state/mcts_state_hypothesis.cpp:62-114)
//   rollout = mamcts RandomHeuristic::calculate_heuristic_values
// (paths relative to /root/reference/bark_mcts/models/behavior/).  The cooperative state
// has its own kernel with the same mapping (rollout_lane_coop_kernel.cuh).
//
// Mapping: T consecutive lanes (T = 1, 2, 4, 8, 16 or 32) own one rollout; a warp carries 32/T
// rollouts and never talks to another warp, so there is no block barrier in the step
// loop.  Within a group the others' IDM hypothesis models are split over the T lanes and
// the ego's single-track macro action runs on lane 0 of the group: every lane of a warp is
// in the same role at the same time (no ego-vs-IDM divergence), and a group that finishes
// its rollout pulls the next one from a global counter, so lanes stay busy until the batch
// is drained (rollouts end at different steps).  Batches too small to fill the GPU run the
// "ego ahead" variant (AHEAD = true, below): a block is one warp of rollouts plus one warp that
// computes the egos' macro actions a pass ahead.
//
// Per-rollout state in shared memory, [field][agent][column] with column = group index in
// the warp (bank = column: conflict free): velocity, arc length s on every lane corridor,
// current corridor + hypothesis id, and the planned (s, v) of the synchronous update; which
// agents are inside which corridor (|d| <= half width) is one bit mask per corridor in
// registers, so the leader search only visits the agents of the own corridor.  The
// Cartesian pose of an IDM agent is a function of (corridor, s) and is rebuilt when needed;
// the ego pose lives in registers.  Lane geometry, road polygon, goal, shapes, hypothesis
// table, macro actions and the config are staged once per block.
//
// Every float operation mirrors oracle/bmcts_oracle.c operation by operation (compile with
// --fmad=false; FMAs only where bm_fma is written).
#pragma once

#include <cstddef>

#include "rollout_device.cuh"

namespace bmcts {

// per-agent state fields of a column
enum { L_V = 0, L_NS, L_NV, L_CUR, L_PROJ };  // L_PROJ + c = arc length s on corridor c

constexpr size_t kScnPrefixBytes = offsetof(bmcts_scenario, hypotheses);
static_assert(kScnPrefixBytes % 8 == 0, "scenario prefix is staged with 8-byte copies");
static_assert(sizeof(bmcts_hypothesis) % 16 == 0, "hypotheses are staged with 16-byte copies");

struct LaneLayout {
  size_t off_acts, off_cfg, off_segcs, off_segsn, off_hflags, off_rad, off_state, warp_bytes, total;
};

__host__ __device__ inline LaneLayout lane_layout(int A, int C, int H, int T, int warps) {
  LaneLayout l;
  size_t b = smem_align16(kScnPrefixBytes + sizeof(bmcts_hypothesis) * (size_t)(H > 0 ? H : 1));
  l.off_acts = b;
  b += smem_align16(sizeof(bmcts_ego_action) * BMCTS_MAX_EGO_ACTIONS);
  l.off_cfg = b;
  b += smem_align16(sizeof(bmcts_config));
  l.off_segcs = b;
  b += sizeof(float) * BMCTS_MAX_CORRIDORS * BMCTS_MAX_CORRIDOR_PTS;
  l.off_segsn = b;
  b += sizeof(float) * BMCTS_MAX_CORRIDORS * BMCTS_MAX_CORRIDOR_PTS;
  l.off_hflags = b;
  b += smem_align16(sizeof(uint32_t) * 4 * (size_t)(H > 0 ? H : 1));  // flags, 1/v0, 1/(2 sqrt(ab))
  l.off_rad = b;  // per agent: plain / static-inflated bounding radius; then the goal's AABB
  // + goal AABB + 2 inner rects + safe arc intervals [lo | hi][corridor][segment]
  // (+ the AABBs of the other agents' goals, cooperative state)
  b += smem_align16(sizeof(float) * (2 * BMCTS_MAX_AGENTS + 4 + 8 +
                                     2 * BMCTS_MAX_CORRIDORS * BMCTS_MAX_CORRIDOR_PTS +
                                     4 * BMCTS_MAX_EXTRA_GOALS));
  l.off_state = b;
  l.warp_bytes = sizeof(float) * (size_t)(L_PROJ + C) * A * (32 / T);
  l.total = b + l.warp_bytes * warps;
  return l;
}

template <int T>
__device__ __forceinline__ uint32_t group_or(uint32_t x) {
#pragma unroll
  for (int o = 1; o < T; o <<= 1) x |= __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

template <int T>
__device__ __forceinline__ float group_lead(float x, int lane) {
  return T == 1 ? x : __shfl_sync(0xffffffffu, x, lane & ~(T - 1));
}

// The work counter resets itself: counter[0] hands out rollouts, counter[1] counts the warps
// that are out of work; the last one of the grid clears both, so the next launch on the stream
// finds them zero and no memset travels with a launch (a driver call per MCTS wave less).  That
// last warp also raises the wave's completion flag in host memory, when there is one: every
// warp's results are ordered before its count (fence), the count before the flag.
__device__ __forceinline__ void retire_warp(const KernelArgs& p, unsigned int* counter,
                                            unsigned int warps_in_grid, int lane) {
  __syncwarp();
  if (lane == 0) {
    if (p.done_flag) __threadfence_system(); else __threadfence();
    if (atomicAdd(counter + 1, 1u) == warps_in_grid - 1u) {
      counter[0] = 0u;
      counter[1] = 0u;
      if (p.done_flag) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned int*>(p.done_flag) = p.done_value;
      }
    }
  }
}

template <int T>
__device__ __forceinline__ void group_sync() {
  if (T > 1) __syncwarp();
}

// Conservative reach of an agent's rectangle around its reference point.  Two rectangles whose
// reference points are farther apart than the sum of their reaches (+1 cm) in x or in y are
// separated on one of the first rectangle's own axes with a margin far above FP32 rounding, so
// the separating-axis test is skipped for such pairs without changing any result:
// |t| > sqrt(2) (hlA + hwA + hlB + hwB) puts max(|t.a1|, |t.a2|) above every SAT threshold.
__device__ __forceinline__ void stage_reach(const bmcts_scenario* scn, const bmcts_config* cfg,
                                            int A, float* rad, int tid, int nthreads) {
  for (int a = tid; a < A; a += nthreads) {
    const float hl = 0.5f * scn->agent_length[a], hw = 0.5f * scn->agent_width[a];
    const float off = bm_abs(0.5f * (scn->agent_length[a] - 2.0f * scn->agent_rear[a]));
    const float plain = 1.42f * (hl + hw) + off + 0.005f;
    rad[a] = plain;
    rad[BMCTS_MAX_AGENTS + a] =
        plain + (cfg->add_safe_dist ? 1.42f * (cfg->static_lat_safe_dist + cfg->static_lon_safe_dist)
                                    : 0.0f);
  }
  if (tid == 0) {  // goal polygon AABB, grown by 1 cm: a point outside it is outside the goal
    float* bb = rad + 2 * BMCTS_MAX_AGENTS;
    float* bb_extra = bb + 12 + 2 * BMCTS_MAX_CORRIDORS * BMCTS_MAX_CORRIDOR_PTS;
    for (int k = 0; k <= BMCTS_MAX_EXTRA_GOALS; ++k) {
      const bmcts_goal& gl = k == 0 ? scn->goal : scn->extra_goals[k - 1];
      float x0 = kBig, x1 = -kBig, y0 = kBig, y1 = -kBig;
      for (int i = 0; i < gl.n_pts && i < BMCTS_MAX_GOAL_PTS; ++i) {
        x0 = bm_min(x0, gl.px[i]);
        x1 = bm_max(x1, gl.px[i]);
        y0 = bm_min(y0, gl.py[i]);
        y1 = bm_max(y1, gl.py[i]);
      }
      float* o = k == 0 ? bb : bb_extra + 4 * (k - 1);
      o[0] = x0 - 0.01f;
      o[1] = x1 + 0.01f;
      o[2] = y0 - 0.01f;
      o[3] = y1 + 0.01f;
    }
    // inner rectangles of the road polygon (engine.cu stores them behind the scenario)
    const float* rc = reinterpret_cast<const float*>(scn + 1);
    for (int i = 0; i < 8; ++i) bb[4 + i] = rc[i];
  }
  // safe arc intervals of the centre lines (behind the rectangles: 8 floats + n + 3 pad words)
  {
    const float* arcs = reinterpret_cast<const float*>(scn + 1) + 12;
    float* dst = rad + 2 * BMCTS_MAX_AGENTS + 12;
    for (int i = tid; i < 2 * BMCTS_MAX_CORRIDORS * BMCTS_MAX_CORRIDOR_PTS; i += nthreads)
      dst[i] = arcs[i];
  }
}

// inside one of the road polygon's inner rectangles (inner_rects.hpp): certainly drivable
__device__ __forceinline__ bool in_rects(const float* rc, float x, float y) {
  const bool a = x >= rc[0] && x <= rc[1] && y >= rc[2] && y <= rc[3];
  const bool b = x >= rc[4] && x <= rc[5] && y >= rc[6] && y <= rc[7];
  return a || b;
}

// IDM sub-step loop (BaseIDM::Plan, spec item 4): returns the travelled distance
#ifndef BMCTS_EGO_UNROLL
#define BMCTS_EGO_UNROLL 1
#endif
#ifndef BMCTS_IDM_UNROLL
#define BMCTS_IDM_UNROLL 2  // tools/sweep_unroll.sh (profiles/README.md): 2 / 1 / 1 is the fastest
#endif

template <bool EXP4>
__device__ __forceinline__ float idm_substeps(const IdmLimits& hl, const IdmParams& ip,
                                              float inv_v0, float inv_2sab, bool has_leader,
                                              float gap, float v_l, float dt, int nsub, float& nv,
                                              float& acc_out) {
  const float half_dt2 = (0.5f * dt) * dt;
  const uint32_t lead_mask = has_leader ? 0xffffffffu : 0u;
  float trav = 0.0f, acc = 0.0f;
BM_UNROLL(BMCTS_IDM_UNROLL)
  for (int q = 0; q < nsub; ++q) {
    const float xr = nv * inv_v0;
    const float pw = EXP4 ? ((xr * xr) * xr) * xr : bm_powi(xr, hl.exponent);
    const float free_term = 1.0f - pw;
    const float dv = nv - v_l;
    const float s_star = ip.s0 + bm_max(0.0f, bm_fma(nv, ip.T, (nv * dv) * inv_2sab));
    // without a leader the interaction term is exactly zero: free_term - 0*0 == free_term;
    // lead_mask is all ones / all zeros, so the quotient is kept or cleared without a branch
    const float qd = bm_div(s_star, bm_max(gap, kGapEps));
    const float qq = __uint_as_float(__float_as_uint(qd) & lead_mask);
    acc = bm_clamp(ip.a_max * (free_term - qq * qq), hl.acc_lo, hl.acc_hi);
    const float ds = bm_max(0.0f, bm_fma(acc, half_dt2, nv * dt));
    nv = bm_clamp(bm_fma(acc, dt, nv), hl.v_min, hl.v_max);
    gap = gap + bm_fma(v_l, dt, -ds);
    trav = trav + ds;
  }
  acc_out = acc;
  return trav;
}

// The ego's macro action (BehaviorMPMacroActions primitive on the single-track model, spec
// item 3): advances the pose (x, y, th, v) and the sine / cosine of the heading over Dt from
// the corridor `cur` the ego is on; returns the clamped acceleration.
__device__ __forceinline__ float ego_macro_step(const bmcts_scenario& scn,
                                                const bmcts_ego_action* acts,
                                                const bmcts_config& cfg, int action, int cur,
                                                float Dt, float& ex, float& ey, float& eth,
                                                float& ev, float& esn, float& ecs) {
  const bmcts_ego_action act = acts[action];
  const float acc = bm_clamp(act.acceleration, cfg.lon_acc_min, cfg.lon_acc_max);
  int tgt = cur;
  if (act.type == BMCTS_ACT_CHANGE_LEFT && scn.corridors[tgt].left >= 0)
    tgt = scn.corridors[tgt].left;
  else if (act.type == BMCTS_ACT_CHANGE_RIGHT && scn.corridors[tgt].right >= 0)
    tgt = scn.corridors[tgt].right;
  const bmcts_corridor& corr = scn.corridors[tgt];
  const float wb = cfg.wheel_base, kct = cfg.crosstrack_gain, dmax = cfg.delta_max;
  const float vmax = cfg.ego_max_velocity;
  const float inv_wb = bm_div(1.0f, wb);
  const int nsub_e = ego_substeps(cfg, Dt);
  const float dt_e = bm_div(Dt, (float)nsub_e);
  const bool limit = cfg.limit_steering_rate != 0;
  const float lat_lo = cfg.lat_acc_min, lat_hi = cfg.lat_acc_max;
  float nx = ex, ny = ey, nth = eth, nv = ev;
  float sn = esn, cs = ecs;  // sin / cos of the current heading (published values)
BM_UNROLL(BMCTS_EGO_UNROLL)
  for (int q = 0; q < nsub_e; ++q) {
    const float fx = bm_fma(wb, cs, nx);
    const float fy = bm_fma(wb, sn, ny);
    const Proj pr = project(corr, fx, fy);
    const float th_err = bm_norm_angle(corr.angle[pr.seg] - nth);
    float delta = th_err + bm_atan2(-(kct * pr.d), nv);
    delta = bm_clamp(delta, -dmax, dmax);
    float tn = bm_tan_small(delta);
    if (limit) {  // lateral acceleration v^2 tan(delta) / L inside [LatAccMin, LatAccMax]
      const float r = bm_div(wb, bm_max(nv * nv, 1.0e-6f));
      tn = bm_clamp(tn, lat_lo * r, lat_hi * r);
    }
    const float dv = dt_e * nv;
    nx = bm_fma(dv, cs, nx);
    ny = bm_fma(dv, sn, ny);
    nth = bm_fma(dv, tn * inv_wb, nth);
    nv = bm_clamp(bm_fma(dt_e, acc, nv), 0.0f, vmax);
    // heading for the next sub-step; after the last one: of the wrapped final heading
    if (q + 1 == nsub_e) nth = bm_norm_angle(nth);
    bm_sincos(nth, &sn, &cs);
  }
  ex = nx;
  ey = ny;
  eth = nth;
  ev = nv;
  esn = sn;
  ecs = cs;
  return acc;
}

// ---- "ego ahead" variant (AHEAD = true), for batches too small to fill the GPU ----------
// The ego's motion does not depend on the other agents (its action is drawn from the
// counter-based generator, its primitive tracks a lane), so the serial chain of its macro
// action -- half of a step's latency when a rollout has a lane per agent -- is taken off the
// step: the block gets a second warp whose lane c computes the ego of column c ONE PASS AHEAD
// of the warp that carries the rollouts and hands it over through shared memory.  The two warps
// meet at one named barrier per pass (no spinning).  Same arithmetic in the same order: results
// are bit-identical to the plain variant.
struct EgoAcq {      // the rollout a column works on, posted by the rollout warp every pass
  uint32_t seq;      // changes with every acquisition
  uint32_t rid;
  int depth;
  uint32_t flags;    // bit 0: expansion step first, bit 1: ego alive in the leaf state
};
struct EgoMail {     // everything is double-buffered by pass parity: the writer of pass i + 1
                     // never touches what the other warp reads in pass i
  uint32_t done[4];  // [parity]: all columns of the rollout warp are out of work
  EgoAcq acq[2][32];
  float slot[2][32][8];  // [parity][column]: x, y, th, v, cos, sin after the pass's ego step
};

__device__ __forceinline__ void ahead_barrier() { asm volatile("bar.sync 1, 64;" ::: "memory"); }

#ifndef BMCTS_LANE_MAX_THREADS
#define BMCTS_LANE_MAX_THREADS 384
#endif
#ifndef BMCTS_LANE_MIN_BLOCKS
#define BMCTS_LANE_MIN_BLOCKS 2
#endif

template <int T, bool AHEAD>
__global__ void __launch_bounds__(AHEAD ? 64 : BMCTS_LANE_MAX_THREADS, AHEAD ? 1 : BMCTS_LANE_MIN_BLOCKS)
rollout_lane_kernel(const KernelArgs p, unsigned int* __restrict__ counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NC = 32 / T;  // rollouts (columns) per warp
  const int A = p.A, C = p.C;
  // AHEAD: one warp of rollouts + the ego warp (launched with 64 threads)
  const int warps = AHEAD ? 1 : blockDim.x >> 5;
  const LaneLayout lay = lane_layout(A, C, p.H, T, warps);

  // ---- stage scenario prefix + used hypotheses, macro actions, config, segment sin/cos,
  //      per-hypothesis "which Philox blocks are needed" flags
  {
    const size_t nb = kScnPrefixBytes + sizeof(bmcts_hypothesis) * (size_t)p.H;
    const uint2* src = reinterpret_cast<const uint2*>(p.scn);
    uint2* dst = reinterpret_cast<uint2*>(smem_raw);
    for (int i = threadIdx.x; i < (int)(nb / 8); i += blockDim.x) dst[i] = src[i];
    const uint32_t* s1 = reinterpret_cast<const uint32_t*>(p.scn->ego_actions);
    uint32_t* d1 = reinterpret_cast<uint32_t*>(smem_raw + lay.off_acts);
    for (int i = threadIdx.x; i < (int)(sizeof(bmcts_ego_action) * BMCTS_MAX_EGO_ACTIONS / 4);
         i += blockDim.x)
      d1[i] = s1[i];
    const uint32_t* s2 = reinterpret_cast<const uint32_t*>(p.cfg);
    uint32_t* d2 = reinterpret_cast<uint32_t*>(smem_raw + lay.off_cfg);
    for (int i = threadIdx.x; i < (int)(sizeof(bmcts_config) / 4); i += blockDim.x) d2[i] = s2[i];
    float* cs = reinterpret_cast<float*>(smem_raw + lay.off_segcs);
    float* sn = reinterpret_cast<float*>(smem_raw + lay.off_segsn);
    for (int i = threadIdx.x; i < BMCTS_MAX_CORRIDORS * BMCTS_MAX_CORRIDOR_PTS; i += blockDim.x) {
      const int c = i / BMCTS_MAX_CORRIDOR_PTS, k = i % BMCTS_MAX_CORRIDOR_PTS;
      float s_ = 0.0f, c_ = 1.0f;
      if (c < C) bm_sincos(p.scn->corridors[c].angle[k], &s_, &c_);
      cs[i] = c_;
      sn[i] = s_;
    }
    stage_reach(p.scn, p.cfg, A, reinterpret_cast<float*>(smem_raw + lay.off_rad), threadIdx.x,
                blockDim.x);
    uint32_t* hf = reinterpret_cast<uint32_t*>(smem_raw + lay.off_hflags);
    for (int i = threadIdx.x; i < p.H; i += blockDim.x) {
      const bmcts_hypothesis& h = p.scn->hypotheses[i];
      // a FixedValue parameter (hi == lo) does not depend on its uniform: u * 0 + lo == lo
      // for every u in [0, 1), so a Philox block none of whose words matter is skipped
      uint32_t f = 0;
      if (h.hi[0] != h.lo[0] || h.hi[1] != h.lo[1] || h.hi[2] != h.lo[2] || h.hi[3] != h.lo[3])
        f |= 1u;
      if (h.hi[4] != h.lo[4]) f |= 2u;
      if (h.exponent == 4) f |= 4u;
      // the per-step constants of a hypothesis whose desired velocity / (max acc, comfortable
      // braking) are FixedValues do not depend on the draw: computed once here
      if (h.hi[3] == h.lo[3]) {
        f |= 8u;
        hf[4 * i + 1] = __float_as_uint(bm_div(1.0f, bm_max(bm_fma(0.0f, h.hi[3] - h.lo[3], h.lo[3]), 1.0e-3f)));
      }
      if (h.hi[2] == h.lo[2] && h.hi[4] == h.lo[4]) {
        f |= 16u;
        const float am_ = bm_fma(0.0f, h.hi[2] - h.lo[2], h.lo[2]);
        const float b_ = bm_fma(0.0f, h.hi[4] - h.lo[4], h.lo[4]);
        hf[4 * i + 2] = __float_as_uint(bm_div(1.0f, 2.0f * bm_sqrt(bm_max(am_ * b_, 1.0e-6f))));
      }
      hf[4 * i] = f;
    }
  }
  __syncthreads();  // the only block barrier: staging done
  const bmcts_scenario& scn = *reinterpret_cast<const bmcts_scenario*>(smem_raw);
  const bmcts_ego_action* acts = reinterpret_cast<const bmcts_ego_action*>(smem_raw + lay.off_acts);
  const bmcts_config& cfg = *reinterpret_cast<const bmcts_config*>(smem_raw + lay.off_cfg);
  const float* segcs = reinterpret_cast<const float*>(smem_raw + lay.off_segcs);
  const float* segsn = reinterpret_cast<const float*>(smem_raw + lay.off_segsn);
  const uint32_t* hflags = reinterpret_cast<const uint32_t*>(smem_raw + lay.off_hflags);
  const float* reach = reinterpret_cast<const float*>(smem_raw + lay.off_rad);
  const float* goal_bb = reach + 2 * BMCTS_MAX_AGENTS;
  const float* rects = goal_bb + 4;
  const float* arc_lo = rects + 8;
  const float* arc_hi = arc_lo + BMCTS_MAX_CORRIDORS * BMCTS_MAX_CORRIDOR_PTS;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  EgoMail* mail = reinterpret_cast<EgoMail*>(smem_raw + smem_align16(lay.total));
  if (AHEAD && warp == 1) {
    // ---------------------------------------------------------------- the ego warp
    uint32_t seq = 0, rid = 0;
    int n = 0, depth = 1;
    bool expand = false, alive = false;
    float ex = 0.0f, ey = 0.0f, eth = 0.0f, ev = 0.0f, ecs = 1.0f, esn = 0.0f;
    for (uint32_t it = 0;; ++it) {
      ahead_barrier();  // acquisitions of this pass posted, slots of this pass complete
      if (mail->done[it & 1u]) break;
      if (lane < NC) {
        const EgoAcq q = mail->acq[it & 1u][lane];
        if (q.seq != seq) {
          seq = q.seq;
          rid = q.rid;
          depth = q.depth;
          expand = (q.flags & 1u) != 0;
          alive = (q.flags & 2u) != 0;
          n = 0;
          const float4 q0 = start_pose(p, 0, rid);
          ex = q0.x;
          ey = q0.y;
          eth = q0.z;
          ev = q0.w;
          bm_sincos(eth, &esn, &ecs);
        }
        if (alive) {
          // the corridor the publish phase finds for this pose: arg min |d| over the corridors
          // the pose projects into (first one on ties, corridor 0 when there is none)
          float best = kBig;
          int cur = 0;
#pragma unroll 1
          for (int c = 0; c < C; ++c) {
            const Proj pr = project(scn.corridors[c], ex, ey);
            const float m = pr.in ? bm_abs(pr.d) : kBig;
            const bool better = m < best;
            best = better ? m : best;
            cur = better ? c : cur;
          }
          const float Dt = p.dt_table
                               ? p.dt_table[depth < BMCTS_MAX_DEPTH ? depth : BMCTS_MAX_DEPTH - 1]
                               : cfg.prediction_k;
          int action;
          if (expand) {
            action = p.action[rid];
          } else {
            uint4 rr = draw(p.key0, p.key1, p.first_rollout_id + (uint64_t)rid, 0u,
                            BMCTS_RNG_ROLLOUT_ACTION, (uint32_t)n);
            action = (int)bm_bounded(rr.x, (uint32_t)scn.n_ego_actions);
          }
          const float acc = ego_macro_step(scn, acts, cfg, action, cur, Dt, ex, ey, eth, ev, esn, ecs);
          if (expand && p.last_action) p.last_action[rid] = acc;
          if (expand) expand = false; else n++;
          depth++;
          float* s = mail->slot[(it + 1u) & 1u][lane];
          s[0] = ex;
          s[1] = ey;
          s[2] = eth;
          s[3] = ev;
          s[4] = ecs;
          s[5] = esn;
        }
      }
    }
    return;
  }
  uint32_t pass = 0, acq_seq = 0;  // AHEAD: pass counter (slot parity), acquisitions posted
  const int t = lane % T;  // lane inside the group
  const int g = lane / T;  // column
  float* col = reinterpret_cast<float*>(smem_raw + lay.off_state + lay.warp_bytes * warp) + g;
  const int fstride = A * NC;  // floats between two fields of the same agent
#define ST(f, a) col[(f) * fstride + (a) * NC]
#define STI(f, a) (reinterpret_cast<int*>(col)[(f) * fstride + (a) * NC])

  const uint32_t n_groups = gridDim.x * (uint32_t)warps * NC;
  const uint32_t my_group = (blockIdx.x * (uint32_t)warps + warp) * NC + g;
  const uint32_t gmask = T == 32 ? 0xffffffffu : (((1u << T) - 1u) << (lane & ~(T - 1)));

  const int nsub = cfg.num_substeps;
  const bool add_sd = cfg.add_safe_dist != 0;
  const bool goal_poly = scn.goal.kind == BMCTS_GOAL_POLYGON;
  const bool goal_frenet = scn.goal.kind == BMCTS_GOAL_FRENET_LIMITS;
  const int n_road = scn.n_road_pts;
  const bool want_final = p.final_pose != nullptr && p.do_rollout != 0;

  // ---- per-rollout registers (replicated in the T lanes of a group)
  bool active = false, done = false, first = true, expand = false, fresh = false;
  uint32_t rid = 0, am = 0;
  int n = 0, depth = 1, agent_steps = 0;
  float disc = 0.0f, R = 0.0f, C0 = 0.0f, C1 = 0.0f, L = 0.0f;
  uint32_t f_any = 0, f_last = 0;
  float ex = 0.0f, ey = 0.0f, eth = 0.0f, ev = 0.0f, ecs = 1.0f, esn = 0.0f;
  float goal_d = 0.0f, goal_ang = 0.0f;
  bool goal_in = false;
  uint32_t inc0 = 0, inc1 = 0, inc2 = 0, inc3 = 0;  // agents inside corridor 0..3
#define INC(c) ((c) == 0 ? inc0 : (c) == 1 ? inc1 : (c) == 2 ? inc2 : inc3)

  for (;;) {
    // ------------------------------------------------ acquire the next rollout
    // A new rollout spends its first pass through the loop in `fresh` mode: only the
    // publish phase runs (Frenet table of the leaf state), the step phases are skipped.
    if (!active && !done) {
      uint32_t next;
      if (first) {
        next = my_group;
        first = false;
      } else {
        next = 0;
        if (t == 0) next = n_groups + atomicAdd(counter, 1u);
        if (T > 1) next = __shfl_sync(gmask, next, lane & ~(T - 1));
      }
      if (next >= (uint32_t)p.n || !wait_chunk(p, next)) {
        done = true;
      } else {
        rid = next;
        // row of the start state in the pose table (re-read where needed: not kept in a register)
        const uint32_t sid = p.state_index ? p.state_index[rid] : rid;
        am = p.alive[sid];
        depth = p.depth ? (int)p.depth[sid] : 1;
        // BMCTS_ACTION_NONE: no expansion step, the rollout starts from the given state
        expand = p.do_expand != 0 && p.action[rid] != BMCTS_ACTION_NONE;
        active = expand || p.do_rollout != 0;
        fresh = true;
        n = 0;
        agent_steps = 0;
        disc = cfg.discount_factor;
        R = C0 = C1 = L = 0.0f;
        f_any = f_last = 0;
        inc0 = inc1 = inc2 = inc3 = 0;
        const float4 q0 = start_pose(p, 0, rid);  // ego pose (agent slot 0)
        ex = q0.x;
        ey = q0.y;
        eth = q0.z;
        ev = q0.w;
        bm_sincos(eth, &esn, &ecs);
        acq_seq++;
      }
    }
    {
      const bool all_done = __all_sync(0xffffffffu, done);
      if (AHEAD) {
        if (t == 0) {
          EgoAcq q;
          q.seq = acq_seq;
          q.rid = rid;
          q.depth = depth;
          q.flags = (expand ? 1u : 0u) | ((am & 1u) && active ? 2u : 0u);
          mail->acq[pass & 1u][g] = q;
        }
        if (lane == 0) mail->done[pass & 1u] = all_done ? 1u : 0u;
        ahead_barrier();  // the ego warp sees the acquisitions; its slots for this pass are written
      }
      if (all_done) {
        retire_warp(p, counter, gridDim.x * (uint32_t)warps, lane);
        break;
      }
    }

    const bool stepping = active && !fresh;
    const uint32_t pre = am;
    const float Dt = p.dt_table
                         ? p.dt_table[depth < BMCTS_MAX_DEPTH ? depth : BMCTS_MAX_DEPTH - 1]
                         : cfg.prediction_k;
    const float dt = bm_div(Dt, (float)nsub);
    const uint64_t rollout_id = p.first_rollout_id + (uint64_t)rid;

    // ---------------------------------------------------------------- predict()
    // others: BehaviorIDMStochastic::Plan under the sampled hypothesis, against the
    // pre-step snapshot; the planned (s, v) are committed in the publish phase
    if (stepping) {
#pragma unroll 1
      for (int i = (t == 0 ? T : t); i < A; i += T) {
        if (!((pre >> i) & 1u)) continue;
        float* ca = col + i * NC;  // this agent's column entries
        const int word = __float_as_int(ca[L_CUR * fstride]);
        const int cur = word & 0xf;
        const int hyp = word >> 8;
        const bmcts_hypothesis& h = scn.hypotheses[hyp];
        const uint32_t hf = hflags[4 * hyp];
        const float* tab_s = col + (L_PROJ + cur) * fstride;
        const float s_cur = tab_s[i * NC];
        // BaseIDM::CalcRelativeValues: nearest alive agent ahead inside the corridor
        int jf = -1;
        float ds_f = kBig;
        if (AHEAD) {
          // latency-bound variant: four candidates per trip, their loads in flight together
          // (a short tail repeats the first candidate, which changes nothing: the test is strict)
          for (uint32_t m = pre & INC(cur) & ~(1u << i); m;) {
            const uint32_t m1 = m & (m - 1u), m2 = m1 & (m1 - 1u), m3 = m2 & (m2 - 1u);
            const int j0 = __ffs(m) - 1;
            const int j1 = m1 ? __ffs(m1) - 1 : j0;
            const int j2 = m2 ? __ffs(m2) - 1 : j0;
            const int j3 = m3 ? __ffs(m3) - 1 : j0;
            m = m3 & (m3 - 1u);
            const float d0 = tab_s[j0 * NC] - s_cur, d1 = tab_s[j1 * NC] - s_cur;
            const float d2 = tab_s[j2 * NC] - s_cur, d3 = tab_s[j3 * NC] - s_cur;
            bool b = d0 > 0.0f && d0 < ds_f;
            ds_f = b ? d0 : ds_f;
            jf = b ? j0 : jf;
            b = d1 > 0.0f && d1 < ds_f;
            ds_f = b ? d1 : ds_f;
            jf = b ? j1 : jf;
            b = d2 > 0.0f && d2 < ds_f;
            ds_f = b ? d2 : ds_f;
            jf = b ? j2 : jf;
            b = d3 > 0.0f && d3 < ds_f;
            ds_f = b ? d3 : ds_f;
            jf = b ? j3 : jf;
          }
        } else {
#pragma unroll 1
          for (uint32_t m = pre & INC(cur) & ~(1u << i); m; m &= m - 1u) {
            const int j = __ffs(m) - 1;
            const float ds = tab_s[j * NC] - s_cur;
            const bool better = ds > 0.0f && ds < ds_f;
            ds_f = better ? ds : ds_f;
            jf = better ? j : jf;
          }
        }
        const bool has_leader = jf >= 0;
        float gap = 0.0f, v_l = 0.0f;
        if (has_leader) {
          gap = ds_f - ((scn.agent_length[i] - scn.agent_rear[i]) + scn.agent_rear[jf]);
          v_l = col[L_V * fstride + jf * NC];
        }
        uint4 r0 = make_uint4(0, 0, 0, 0), r1 = make_uint4(0, 0, 0, 0);
        {
          uint64_t id = rollout_id;
          uint32_t dom = BMCTS_RNG_ROLLOUT_PARAMS, idx = 2u * (uint32_t)n;
          if (expand) {
            id = p.draw_id ? p.draw_id[(size_t)i * p.n + rid] : 0ull;
            dom = BMCTS_RNG_TREE_PARAMS;
            idx = 0u;
          }
          if (hf & 1u) r0 = draw(p.key0, p.key1, id, (uint32_t)i, dom, idx);
          if (hf & 2u) r1 = draw(p.key0, p.key1, id, (uint32_t)i, dom, idx + 1u);
        }
        const IdmParams ip = sample_idm_params(h, r0, r1);
        const IdmLimits hl = idm_limits(h);
        const float inv_v0 = (hf & 8u) ? __uint_as_float(hflags[4 * hyp + 1])
                                       : bm_div(1.0f, bm_max(ip.v0, 1.0e-3f));
        const float inv_2sab =
            (hf & 16u) ? __uint_as_float(hflags[4 * hyp + 2])
                       : bm_div(1.0f, 2.0f * bm_sqrt(bm_max(ip.a_max * ip.b, 1.0e-6f)));
        float nv = ca[L_V * fstride], acc;
        const float trav =
            (hf & 4u) ? idm_substeps<true>(hl, ip, inv_v0, inv_2sab, has_leader, gap, v_l, dt, nsub,
                                           nv, acc)
                      : idm_substeps<false>(hl, ip, inv_v0, inv_2sab, has_leader, gap, v_l, dt, nsub,
                                            nv, acc);
        ca[L_NS * fstride] = bm_min(s_cur + trav, scn.corridors[cur].length);
        ca[L_NV * fstride] = nv;
        if (expand && p.last_action) p.last_action[(size_t)i * p.n + rid] = acc;
      }
    }

    // ego: BehaviorMPMacroActions primitive on the single-track model (lane 0 of the group;
    // AHEAD: computed one pass ago by the ego warp)
    if (AHEAD) {
      if (stepping && (pre & 1u)) {
        const float* s = mail->slot[pass & 1u][g];
        ex = s[0];
        ey = s[1];
        eth = s[2];
        ev = s[3];
        ecs = s[4];
        esn = s[5];
      } else if (stepping && expand && t == 0 && p.last_action) {
        p.last_action[rid] = 0.0f;
      }
    } else {
      if (stepping && (pre & 1u) && t == 0) {
        int action;
        if (expand) {
          action = p.action[rid];
        } else {
          uint4 rr = draw(p.key0, p.key1, rollout_id, 0u, BMCTS_RNG_ROLLOUT_ACTION, (uint32_t)n);
          action = (int)bm_bounded(rr.x, (uint32_t)scn.n_ego_actions);
        }
        const float acc =
            ego_macro_step(scn, acts, cfg, action, STI(L_CUR, 0) & 0xf, Dt, ex, ey, eth, ev, esn, ecs);
        if (expand && p.last_action) p.last_action[rid] = acc;
      } else if (stepping && expand && t == 0 && p.last_action) {
        p.last_action[rid] = 0.0f;
      }
      if (T > 1) {
        ex = group_lead<T>(ex, lane);
        ey = group_lead<T>(ey, lane);
        eth = group_lead<T>(eth, lane);
        ev = group_lead<T>(ev, lane);
        ecs = group_lead<T>(ecs, lane);
        esn = group_lead<T>(esn, lane);
      }
    }
    group_sync<T>();  // every read of the pre-step Frenet table is done

    // ------------------------------------------------------------ publish phase
    // Every agent of the group: new pose (ego: registers; others: centre-line point of the
    // planned s; fresh rollout: the leaf pose), World::RemoveInvalidAgents, Frenet table of
    // the post-step world, and the pair tests against the ego's post-step rectangle.
    const Box be = agent_box(scn, 0, ex, ey, ecs, esn, 0.0f, 0.0f);
    const Box bs = agent_box(scn, 0, ex, ey, ecs, esn, cfg.static_lat_safe_dist,
                             cfg.static_lon_safe_dist);
    uint32_t fl = 0, dead = 0, pubs = 0, near = 0, in0 = 0, in1 = 0, in2 = 0, in3 = 0;
    if (active) {
#pragma unroll 1
      for (int a = t; a < A; a += T) {
        const size_t gi = (size_t)a * p.n + rid;
        float* ca = col + a * NC;
        const bool was_alive = (pre >> a) & 1u;
        float x, y, th, v;
        int word = 0;
        bool on_safe_arc = false;  // centre-line point with a margin to the road boundary
        if (fresh) {
          const float4 q = start_pose(p, a, rid);
          x = q.x;
          y = q.y;
          th = q.z;
          v = q.w;
          word = (p.hyp ? (int)p.hyp[gi] : 0) << 8;
          if (want_final) p.final_pose[gi] = q;
          if (!was_alive) ca[L_CUR * fstride] = __int_as_float(word);
        } else if (!was_alive) {
          if (expand) {
            if (p.next_pose) p.next_pose[gi] = start_pose(p, a, rid);
            if (p.reward && a > 0) p.reward[gi] = 0.0f;
            if (p.last_action && a > 0) p.last_action[gi] = 0.0f;
          }
        } else if (a == 0) {
          x = ex;
          y = ey;
          th = eth;
          v = ev;
        } else {
          word = __float_as_int(ca[L_CUR * fstride]);
          const int cur = word & 0xf;
          const bmcts_corridor& corr = scn.corridors[cur];
          const float s_new = ca[L_NS * fstride];
          const int nseg = corr.n_pts - 1;
          int sg = 0;
#pragma unroll 1
          for (int q = 1; q < nseg; ++q) sg = (corr.s0[q] <= s_new) ? q : sg;
          const float ls = s_new - corr.s0[sg];
          x = bm_fma(ls, corr.tx[sg], corr.px[sg]);
          y = bm_fma(ls, corr.ty[sg], corr.py[sg]);
          th = corr.angle[sg];
          v = ca[L_NV * fstride];
          on_safe_arc = ls >= arc_lo[cur * BMCTS_MAX_CORRIDOR_PTS + sg] &&
                        ls <= arc_hi[cur * BMCTS_MAX_CORRIDOR_PTS + sg];
          word = (word & ~0xff) | (cur << 4);  // bits 4-7: the corridor the pose was built on
        }
        if (!was_alive) continue;
        if (!fresh) {
          if (expand) {
            if (p.next_pose) p.next_pose[gi] = make_float4(x, y, th, v);
            if (p.reward && a > 0) p.reward[gi] = 0.0f;
          }
          if (want_final) p.final_pose[gi] = make_float4(x, y, th, v);
          // World::RemoveInvalidAgents; a centre-line point on a safe arc or a point inside an
          // inner rectangle of the road polygon needs no scan
          if (!on_safe_arc && !in_rects(rects, x, y) &&
              !point_in_polygon(scn.road_x, scn.road_y, n_road, x, y)) {
            dead |= 1u << a;
            continue;
          }
        }
        // Frenet projection on every lane corridor; current corridor = arg min |d|
        float best = kBig;
        int bc = 0;
        const uint32_t bit = 1u << a;
        pubs |= bit;
#pragma unroll 1
        for (int c = 0; c < C; ++c) {
          const bmcts_corridor& cc = scn.corridors[c];
          const Proj pr = project(cc, x, y);
          const float ad = bm_abs(pr.d);
          const float m = pr.in ? ad : kBig;
          ca[(L_PROJ + c) * fstride] = pr.s;
          const uint32_t inb = (pr.in && ad <= cc.half_width) ? bit : 0u;
          if (c == 0) in0 |= inb;
          if (c == 1) in1 |= inb;
          if (c == 2) in2 |= inb;
          if (c == 3) in3 |= inb;
          const bool better = m < best;
          best = better ? m : best;
          bc = better ? c : bc;
          if (goal_frenet && a == 0 && c == scn.goal.corridor) {
            goal_d = pr.d;
            goal_in = pr.in;
            goal_ang = bm_norm_angle(th - cc.angle[pr.seg]);
          }
        }
        ca[L_CUR * fstride] = __int_as_float(word | bc);
        ca[L_V * fstride] = v;
        if (!fresh && a > 0) {
          // pair (ego, a): only agents within reach of the ego's (inflated) rectangle need
          // the separating-axis tests; they are done in a second, compacted loop
          const float thr = reach[BMCTS_MAX_AGENTS] + reach[a];
          if (bm_abs(x - ex) <= thr && bm_abs(y - ey) <= thr) near |= bit;
        }
      }
      // EvaluatorCollisionEgoAgent / EvaluatorStaticSafeDist for the agents in reach; the pose
      // is rebuilt from the planned s on the corridor it was built on.  Masked below when the
      // ego itself left the map.
#pragma unroll 1
      for (uint32_t m = near; m; m &= m - 1u) {
        const int a = __ffs(m) - 1;
        const float* ca = col + a * NC;
        const int oc = (__float_as_int(ca[L_CUR * fstride]) >> 4) & 0xf;
        const bmcts_corridor& corr = scn.corridors[oc];
        const float s_new = ca[L_NS * fstride];
        const int nseg = corr.n_pts - 1;
        int sg = 0;
#pragma unroll 1
        for (int q = 1; q < nseg; ++q) sg = (corr.s0[q] <= s_new) ? q : sg;
        const float ls = s_new - corr.s0[sg];
        const float x = bm_fma(ls, corr.tx[sg], corr.px[sg]);
        const float y = bm_fma(ls, corr.ty[sg], corr.py[sg]);
        const Box bj = agent_box(scn, a, x, y, segcs[oc * BMCTS_MAX_CORRIDOR_PTS + sg],
                                 segsn[oc * BMCTS_MAX_CORRIDOR_PTS + sg], 0.0f, 0.0f);
        if (boxes_overlap(be, bj)) fl |= BMCTS_F_COLLISION_OTHER;
        if (add_sd && boxes_overlap(bs, bj)) fl |= BMCTS_F_STATIC_SAFE_DIST;
      }
    }
    if (T > 1) {
      dead = group_or<T>(dead);
      fl = group_or<T>(fl);
      pubs = group_or<T>(pubs);
      in0 = group_or<T>(in0);
      if (C > 1) in1 = group_or<T>(in1);
      if (C > 2) in2 = group_or<T>(in2);
      if (C > 3) in3 = group_or<T>(in3);
      goal_d = group_lead<T>(goal_d, lane);
      goal_ang = group_lead<T>(goal_ang, lane);
      goal_in = __shfl_sync(0xffffffffu, (int)goal_in, lane & ~(T - 1)) != 0;
    }
    am &= ~dead;
    inc0 = (inc0 & ~pubs) | in0;
    inc1 = (inc1 & ~pubs) | in1;
    inc2 = (inc2 & ~pubs) | in2;
    inc3 = (inc3 & ~pubs) | in3;
    const bool ego_alive = (am & 1u) != 0;
    group_sync<T>();  // post-step Frenet table complete

    // --------------------------------------- mcts_observed_world_evaluation (ego view)
    float reward = 0.0f, c0 = 0.0f, c1 = 0.0f;
    uint32_t f = ego_alive ? fl : 0u;
    if (stepping && ego_alive) {
      // EvaluatorSafeDistDrivableArea / EvaluatorGoalReached: corner-in-polygon tests,
      // items 0-3 = inflated shape vs road corridor, 4-7 = shape vs goal polygon, split
      // over the lanes of the group
      const Box bd = agent_box(scn, 0, ex, ey, ecs, esn, cfg.drivable_lat_safe_dist,
                               cfg.drivable_lon_safe_dist);
      const int n_items = goal_poly ? 8 : 4;
#pragma unroll 1
      for (int q = t; q < n_items; q += T) {
        const bool road = q < 4;
        if (!road && (f & kGoalFailBit)) continue;  // one corner outside: goal not reached
        float cx, cy;
        box_corner(road ? bd : be, q & 3, cx, cy);
        bool in;
        if (!road && (cx < goal_bb[0] || cx > goal_bb[1] || cy < goal_bb[2] || cy > goal_bb[3]))
          in = false;  // outside the goal's bounding box (1 cm margin)
        else if (road && in_rects(rects, cx, cy))
          in = true;  // inside an inner rectangle of the road polygon
        else
          in = point_in_polygon(road ? scn.road_x : scn.goal.px, road ? scn.road_y : scn.goal.py,
                                road ? n_road : scn.goal.n_pts, cx, cy);
        if (!in) f |= road ? (uint32_t)BMCTS_F_COLLISION_DRIVABLE : kGoalFailBit;
      }
    }
    if (T > 1) f = group_or<T>(f);
    if (stepping) {
      if (!ego_alive) {
        f |= BMCTS_F_OUT_OF_MAP;
      } else {
        if (goal_poly && !(f & kGoalFailBit)) f |= BMCTS_F_GOAL_REACHED;
        if (goal_frenet) {
          const bmcts_goal& gl = scn.goal;
          bool ok = goal_in && goal_d <= gl.max_lat_left && -goal_d <= gl.max_lat_right &&
                    goal_ang <= gl.max_angle_left && -goal_ang <= gl.max_angle_right &&
                    ev >= gl.vel_lo && ev <= gl.vel_hi;
          if (ok) f |= BMCTS_F_GOAL_REACHED;
        }
        if (add_sd) {
          // EvaluatorDynamicSafeDist: nearest agents ahead of / behind the ego in its corridor
          const int cur = STI(L_CUR, 0) & 0xf;
          const float* tab_s = col + (L_PROJ + cur) * fstride;
          const float s_e = tab_s[0];
          int jf = -1, jr = -1;
          float ds_f = kBig, ds_r = kBig;
          if (AHEAD) {  // (two candidates per trip, as in the leader search above)
            for (uint32_t m = am & INC(cur) & ~1u; m;) {
              const uint32_t m1 = m & (m - 1u);
              const int j0 = __ffs(m) - 1;
              const int j1 = m1 ? __ffs(m1) - 1 : j0;
              m = m1 & (m1 - 1u);
              const float d0 = tab_s[j0 * NC] - s_e, d1 = tab_s[j1 * NC] - s_e;
              bool bf = d0 > 0.0f && d0 < ds_f, br = -d0 > 0.0f && -d0 < ds_r;
              ds_f = bf ? d0 : ds_f;
              jf = bf ? j0 : jf;
              ds_r = br ? -d0 : ds_r;
              jr = br ? j0 : jr;
              bf = d1 > 0.0f && d1 < ds_f;
              br = -d1 > 0.0f && -d1 < ds_r;
              ds_f = bf ? d1 : ds_f;
              jf = bf ? j1 : jf;
              ds_r = br ? -d1 : ds_r;
              jr = br ? j1 : jr;
            }
          } else {
#pragma unroll 1
            for (uint32_t m = am & INC(cur) & ~1u; m; m &= m - 1u) {
              const int j = __ffs(m) - 1;
              const float ds = tab_s[j * NC] - s_e;
              const float dr = -ds;
              const bool bf = ds > 0.0f && ds < ds_f;
              const bool br = dr > 0.0f && dr < ds_r;
              ds_f = bf ? ds : ds_f;
              jf = bf ? j : jf;
              ds_r = br ? dr : ds_r;
              jr = br ? j : jr;
            }
          }
          bool safe = true;
          if (jf >= 0) {
            float dist = ds_f - ((scn.agent_length[0] - scn.agent_rear[0]) + scn.agent_rear[jf]);
            safe = safe && safe_dist_longitudinal(ST(L_V, jf), ev, dist, -cfg.dyn_max_ego_decel,
                                                  -cfg.dyn_max_other_decel,
                                                  cfg.dyn_reaction_time_ego);
          }
          if (jr >= 0 && cfg.dyn_to_rear) {
            float dist = ds_r - ((scn.agent_length[jr] - scn.agent_rear[jr]) + scn.agent_rear[0]);
            safe = safe && safe_dist_longitudinal(ev, ST(L_V, jr), dist, -cfg.dyn_max_other_decel,
                                                  -cfg.dyn_max_ego_decel,
                                                  cfg.dyn_reaction_time_others);
          }
          if (!safe) f |= BMCTS_F_DYNAMIC_SAFE_DIST;
        }
      }
      f &= 0xffu;
      f |= terminal_from_flags(cfg, f);
      // generate_next_state (mcts_state_hypothesis.hpp:125-140, risk_constraint.cpp:67-81)
      reward = reward_from_flags(cfg, f, Dt);
      costs_from_flags(cfg, f, Dt, c0, c1);
    }

    // ------------------------------------------------ outputs / accumulation
    if (stepping) {
      bool finished = false;
      if (expand) {
        if (t == 0) {
          if (p.reward) p.reward[rid] = reward;
          if (p.next_alive) p.next_alive[rid] = am;
          if (p.cost) {
            p.cost[rid] = c0;
            p.cost[(size_t)p.n + rid] = c1;
          }
          if (p.flags) p.flags[rid] = f;
        }
        if (!p.do_rollout || (f & BMCTS_F_TERMINAL) || cfg.max_rollout_steps <= 0) {
          // RandomHeuristic on a terminal child: zero estimate, flags of the expansion step
          finished = true;
          f_last = f;
          f_any = f;
        }
        expand = false;
      } else {
        agent_steps += __popc(pre);
        R = bm_fma(disc, reward, R);
        C0 = bm_fma(disc, c0, C0);
        C1 = bm_fma(disc, c1, C1);
        L = L + Dt;
        f_last = f;
        f_any |= f;
        if ((f & BMCTS_F_TERMINAL) || n + 1 >= cfg.max_rollout_steps) finished = true;
        disc = disc * cfg.discount_factor;
        n++;
      }
      depth++;
      if (finished) {
        active = false;
        if (p.do_rollout) {
          if (t == 0 && p.result) {
            float4* dst = reinterpret_cast<float4*>(&p.result[rid]);
            dst[0] = make_float4(R, C0, C1, L);
            dst[1] = make_float4(__uint_as_float((f_last & 0xffu) | ((f_any & 0xffu) << 8)),
                                 __int_as_float(n), __int_as_float(agent_steps),
                                 __uint_as_float(am));
          }
          if (p.ret_agents) {
#pragma unroll 1
            for (int a = t; a < A; a += T)
              p.ret_agents[(size_t)a * p.n + rid] = (a == 0) ? R : 0.0f;
          }
        }
      }
    }
    fresh = false;
    pass++;
  }
#undef ST
#undef STI
#undef INC
}

}  // namespace bmcts
