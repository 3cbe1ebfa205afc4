// rollout_lane_kernel.cuh -- sm_100a rollout kernel of the hypothesis / risk-constraint
// states, "rollout-per-lane-group" mapping.
//
// Replaces, for MctsStateHypothesis<void> and MctsStateRiskConstraint,
//   step    = MctsStateHypothesis<T>::execute + plan_action_current_hypothesis
//             (mcts_state/mcts_state_hypothesis.cpp:62-114)
//   rollout = mamcts RandomHeuristic::calculate_heuristic_values
// (paths relative to /root/reference/bark_mcts/models/behavior/).  The cooperative state
// keeps the agent-per-warp kernel of rollout_kernel.cuh.
//
// Mapping: T consecutive lanes (T = 1, 2 or 4) own one rollout; a warp carries 32/T
// rollouts and never talks to another warp, so there is no block barrier in the step
// loop.  Within a group the others' IDM hypothesis models are split over the T lanes and
// the ego's single-track macro action runs on lane 0 of the group: every lane of a warp is
// in the same role at the same time (no ego-vs-IDM divergence), and a group that finishes
// its rollout pulls the next one from a global counter, so lanes stay busy until the batch
// is drained (rollouts end at different steps).
//
// Per-rollout state in shared memory, [field][agent][column] with column = group index in
// the warp (bank = column: conflict free): velocity, Frenet (s, d) on every lane corridor,
// current corridor + hypothesis id, and the planned (s, v) of the synchronous update.  The
// Cartesian pose of an IDM agent is a function of (corridor, s) and is rebuilt when needed;
// the ego pose lives in registers.  Lane geometry, road polygon, goal, shapes, hypothesis
// table, macro actions and the config are staged once per block.
//
// Every float operation mirrors oracle/bmcts_oracle.c operation by operation (compile with
// --fmad=false; FMAs only where bm_fma is written); the results are bit-identical to the
// agent-per-warp kernel.
#pragma once

#include <cstddef>

#include "rollout_kernel.cuh"

namespace bmcts {

// per-agent state fields of a column
enum { L_V = 0, L_NS, L_NV, L_CUR, L_PROJ };  // L_PROJ + 2*c = s, +1 = d_eff on corridor c

constexpr size_t kScnPrefixBytes = offsetof(bmcts_scenario, hypotheses);
static_assert(kScnPrefixBytes % 8 == 0, "scenario prefix is staged with 8-byte copies");
static_assert(sizeof(bmcts_hypothesis) % 16 == 0, "hypotheses are staged with 16-byte copies");

struct LaneLayout {
  size_t off_acts, off_cfg, off_segcs, off_segsn, off_state, warp_bytes, total;
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
  l.off_state = b;
  l.warp_bytes = sizeof(float) * (size_t)(L_PROJ + 2 * C) * A * (32 / T);
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

template <int T>
__device__ __forceinline__ void group_sync() {
  if (T > 1) __syncwarp();
}

// IDM sub-step loop (BaseIDM::Plan, spec item 4): returns the travelled distance
template <bool EXP4>
__device__ __forceinline__ float idm_substeps(const IdmLimits& hl, const IdmParams& ip,
                                              bool has_leader, float gap, float v_l, float dt,
                                              int nsub, float& nv, float& acc_out) {
  const float inv_v0 = bm_div(1.0f, bm_max(ip.v0, 1.0e-3f));
  const float inv_2sab = bm_div(1.0f, 2.0f * bm_sqrt(bm_max(ip.a_max * ip.b, 1.0e-6f)));
  const float half_dt2 = (0.5f * dt) * dt;
  float trav = 0.0f, acc = 0.0f;
  for (int q = 0; q < nsub; ++q) {
    const float xr = nv * inv_v0;
    const float pw = EXP4 ? ((xr * xr) * xr) * xr : bm_powi(xr, hl.exponent);
    const float free_term = 1.0f - pw;
    const float dv = nv - v_l;
    const float s_star = ip.s0 + bm_max(0.0f, bm_fma(nv, ip.T, (nv * dv) * inv_2sab));
    // without a leader the interaction term is exactly zero: free_term - 0*0 == free_term
    const float qq = has_leader ? bm_div(s_star, bm_max(gap, kGapEps)) : 0.0f;
    acc = bm_clamp(ip.a_max * (free_term - qq * qq), hl.acc_lo, hl.acc_hi);
    const float ds = bm_max(0.0f, bm_fma(acc, half_dt2, nv * dt));
    nv = bm_clamp(bm_fma(acc, dt, nv), hl.v_min, hl.v_max);
    gap = gap + bm_fma(v_l, dt, -ds);
    trav = trav + ds;
  }
  acc_out = acc;
  return trav;
}

template <int T>
__global__ void __launch_bounds__(512) rollout_lane_kernel(const KernelArgs p,
                                                            unsigned int* __restrict__ counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NC = 32 / T;  // rollouts (columns) per warp
  const int A = p.A, C = p.C;
  const int warps = blockDim.x >> 5;
  const LaneLayout lay = lane_layout(A, C, p.H, T, warps);

  // ---- stage scenario prefix + used hypotheses, macro actions, config, segment sin/cos
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
  }
  __syncthreads();  // the only block barrier: staging done
  const bmcts_scenario& scn = *reinterpret_cast<const bmcts_scenario*>(smem_raw);
  const bmcts_ego_action* acts = reinterpret_cast<const bmcts_ego_action*>(smem_raw + lay.off_acts);
  const bmcts_config& cfg = *reinterpret_cast<const bmcts_config*>(smem_raw + lay.off_cfg);
  const float* segcs = reinterpret_cast<const float*>(smem_raw + lay.off_segcs);
  const float* segsn = reinterpret_cast<const float*>(smem_raw + lay.off_segsn);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = lane % T;  // lane inside the group
  const int g = lane / T;  // column
  float* col = reinterpret_cast<float*>(smem_raw + lay.off_state + lay.warp_bytes * warp) + g;
#define ST(f, a) col[((f) * A + (a)) * NC]
#define STI(f, a) (reinterpret_cast<int*>(col)[((f) * A + (a)) * NC])

  const uint32_t n_groups = gridDim.x * (uint32_t)warps * NC;
  const uint32_t my_group = (blockIdx.x * (uint32_t)warps + warp) * NC + g;
  const uint32_t gmask = T == 32 ? 0xffffffffu : (((1u << T) - 1u) << (lane & ~(T - 1)));

  const int nsub = cfg.num_substeps;
  const bool add_sd = cfg.add_safe_dist != 0;
  const bool goal_poly = scn.goal.kind == BMCTS_GOAL_POLYGON;
  const bool goal_frenet = scn.goal.kind == BMCTS_GOAL_FRENET_LIMITS;

  // projection of one agent on every corridor; writes the Frenet table, returns cur
  auto publish_agent = [&](int a, float x, float y, float th, bool want_goal, float& goal_d,
                           bool& goal_in, float& goal_ang) -> int {
    float best = kBig;
    int bc = 0;
    for (int c = 0; c < C; ++c) {
      Proj pr = project(scn.corridors[c], x, y);
      float m = pr.in ? bm_abs(pr.d) : kBig;
      ST(L_PROJ + 2 * c, a) = pr.s;
      ST(L_PROJ + 2 * c + 1, a) = pr.in ? pr.d : kBig;
      if (m < best) {
        best = m;
        bc = c;
      }
      if (want_goal && c == scn.goal.corridor) {
        goal_d = pr.d;
        goal_in = pr.in;
        goal_ang = bm_norm_angle(th - scn.corridors[c].angle[pr.seg]);
      }
    }
    return bc;
  };

  // ---- per-rollout registers (replicated in the T lanes of a group)
  bool active = false, done = false, first = true, expand = false;
  uint32_t rid = 0, am = 0;
  int n = 0, depth = 1, agent_steps = 0;
  float disc = 0.0f, R = 0.0f, C0 = 0.0f, C1 = 0.0f, L = 0.0f;
  uint32_t f_any = 0, f_last = 0;
  float ex = 0.0f, ey = 0.0f, eth = 0.0f, ev = 0.0f, ecs = 1.0f, esn = 0.0f;
  float goal_d = 0.0f, goal_ang = 0.0f;
  bool goal_in = false;

  for (;;) {
    // ------------------------------------------------ acquire the next rollout
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
      if (next >= (uint32_t)p.n) {
        done = true;
      } else {
        rid = next;
        am = p.alive[rid];
        depth = p.depth ? (int)p.depth[rid] : 1;
        expand = p.do_expand != 0;
        active = true;
        n = 0;
        agent_steps = 0;
        disc = cfg.discount_factor;
        R = C0 = C1 = L = 0.0f;
        f_any = f_last = 0;
        const float4 q0 = p.pose[rid];  // ego pose (agent slot 0)
        ex = q0.x;
        ey = q0.y;
        eth = q0.z;
        ev = q0.w;
        for (int a = t; a < A; a += T) {
          const size_t gi = (size_t)a * p.n + rid;
          const float4 q = p.pose[gi];
          if (p.final_pose && p.do_rollout) p.final_pose[gi] = q;
          int hyp = p.hyp ? (int)p.hyp[gi] : 0;
          int cur = 0;
          if ((am >> a) & 1u) {
            float gd = 0.0f, ga = 0.0f;
            bool gin = false;
            cur = publish_agent(a, q.x, q.y, q.z, goal_frenet && a == 0, gd, gin, ga);
            ST(L_V, a) = q.w;
          }
          STI(L_CUR, a) = cur | (hyp << 8);
        }
        bm_sincos(eth, &esn, &ecs);
        if (T > 1) __syncwarp(gmask);
      }
    }
    if (__all_sync(0xffffffffu, done)) break;

    // ---------------------------------------------------------------- predict()
    const uint32_t pre = am;
    const float Dt = p.dt_table
                         ? p.dt_table[depth < BMCTS_MAX_DEPTH ? depth : BMCTS_MAX_DEPTH - 1]
                         : cfg.prediction_k;
    const float dt = bm_div(Dt, (float)nsub);
    const uint64_t rollout_id = p.first_rollout_id + (uint64_t)rid;

    // others: BehaviorIDMStochastic::Plan under the sampled hypothesis, against the
    // pre-step snapshot; the planned (s, v) are committed in the second pass
    if (active) {
      for (int i = (t == 0 ? T : t); i < A; i += T) {
        if (!((pre >> i) & 1u)) continue;
        const int word = STI(L_CUR, i);
        const int cur = word & 0xff;
        const bmcts_hypothesis& h = scn.hypotheses[word >> 8];
        const float s_cur = ST(L_PROJ + 2 * cur, i);
        // BaseIDM::CalcRelativeValues: nearest alive agent ahead inside the corridor
        const float hwc = scn.corridors[cur].half_width;
        int jf = -1;
        float ds_f = kBig;
        for (uint32_t m = pre & ~(1u << i); m; m &= m - 1u) {
          const int j = __ffs(m) - 1;
          const float dj = ST(L_PROJ + 2 * cur + 1, j);
          const float ds = ST(L_PROJ + 2 * cur, j) - s_cur;
          if (bm_abs(dj) <= hwc && ds > 0.0f && ds < ds_f) {
            ds_f = ds;
            jf = j;
          }
        }
        const bool has_leader = jf >= 0;
        float gap = 0.0f, v_l = 0.0f;
        if (has_leader) {
          gap = ds_f - ((scn.agent_length[i] - scn.agent_rear[i]) + scn.agent_rear[jf]);
          v_l = ST(L_V, jf);
        }
        // parameter draw; a FixedValue parameter (hi == lo) is independent of the uniform
        // (u * 0 + lo == lo for every u in [0, 1)), so its Philox block is skipped
        const bool need0 = h.hi[0] != h.lo[0] || h.hi[1] != h.lo[1] || h.hi[2] != h.lo[2] ||
                           h.hi[3] != h.lo[3];
        const bool need1 = h.hi[4] != h.lo[4];
        uint4 r0 = make_uint4(0, 0, 0, 0), r1 = make_uint4(0, 0, 0, 0);
        {
          uint64_t id = rollout_id;
          uint32_t dom = BMCTS_RNG_ROLLOUT_PARAMS, idx = 2u * (uint32_t)n;
          if (expand) {
            id = p.draw_id ? p.draw_id[(size_t)i * p.n + rid] : 0ull;
            dom = BMCTS_RNG_TREE_PARAMS;
            idx = 0u;
          }
          if (need0) r0 = draw(p.key0, p.key1, id, (uint32_t)i, dom, idx);
          if (need1) r1 = draw(p.key0, p.key1, id, (uint32_t)i, dom, idx + 1u);
        }
        const IdmParams ip = sample_idm_params(h, r0, r1);
        const IdmLimits hl = idm_limits(h);
        float nv = ST(L_V, i), acc;
        const float trav = (hl.exponent == 4)
                               ? idm_substeps<true>(hl, ip, has_leader, gap, v_l, dt, nsub, nv, acc)
                               : idm_substeps<false>(hl, ip, has_leader, gap, v_l, dt, nsub, nv, acc);
        ST(L_NS, i) = bm_min(s_cur + trav, scn.corridors[cur].length);
        ST(L_NV, i) = nv;
        if (expand && p.last_action) p.last_action[(size_t)i * p.n + rid] = acc;
      }
    }
    group_sync<T>();  // every read of the pre-step Frenet table is done

    // ego: BehaviorMPMacroActions primitive on the single-track model (lane 0 of the group)
    bool ego_alive = (pre & 1u) != 0;
    if (active && ego_alive && t == 0) {
      int action;
      if (expand) {
        action = p.action[rid];
      } else {
        uint4 rr = draw(p.key0, p.key1, rollout_id, 0u, BMCTS_RNG_ROLLOUT_ACTION, (uint32_t)n);
        action = (int)bm_bounded(rr.x, (uint32_t)scn.n_ego_actions);
      }
      const bmcts_ego_action act = acts[action];
      const float acc = bm_clamp(act.acceleration, cfg.lon_acc_min, cfg.lon_acc_max);
      int tgt = STI(L_CUR, 0) & 0xff;
      if (act.type == BMCTS_ACT_CHANGE_LEFT && scn.corridors[tgt].left >= 0)
        tgt = scn.corridors[tgt].left;
      else if (act.type == BMCTS_ACT_CHANGE_RIGHT && scn.corridors[tgt].right >= 0)
        tgt = scn.corridors[tgt].right;
      const bmcts_corridor& corr = scn.corridors[tgt];
      const float inv_wb = bm_div(1.0f, cfg.wheel_base);
      float nx = ex, ny = ey, nth = eth, nv = ev;
      float sn = esn, cs = ecs;  // sin / cos of the current heading (published values)
      for (int q = 0; q < nsub; ++q) {
        if (q > 0) bm_sincos(nth, &sn, &cs);
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
      ex = nx;
      ey = ny;
      eth = bm_norm_angle(nth);
      ev = nv;
      if (expand && p.last_action) p.last_action[rid] = acc;
      // World::RemoveInvalidAgents (spec item 6) and the ego's part of the post-step world
      ego_alive = point_in_polygon(scn.road_x, scn.road_y, scn.n_road_pts, ex, ey);
      if (ego_alive) {
        bm_sincos(eth, &esn, &ecs);
        const int cur = publish_agent(0, ex, ey, eth, goal_frenet, goal_d, goal_in, goal_ang);
        STI(L_CUR, 0) = cur;
        ST(L_V, 0) = ev;
      }
    } else if (active && expand && t == 0 && p.last_action) {
      p.last_action[rid] = 0.0f;
    }
    if (T > 1) {
      ex = group_lead<T>(ex, lane);
      ey = group_lead<T>(ey, lane);
      eth = group_lead<T>(eth, lane);
      ev = group_lead<T>(ev, lane);
      ecs = group_lead<T>(ecs, lane);
      esn = group_lead<T>(esn, lane);
      goal_d = group_lead<T>(goal_d, lane);
      goal_ang = group_lead<T>(goal_ang, lane);
      goal_in = __shfl_sync(0xffffffffu, (int)goal_in, lane & ~(T - 1)) != 0;
      ego_alive = __shfl_sync(0xffffffffu, (int)ego_alive, lane & ~(T - 1)) != 0;
    }
    if (!ego_alive) am &= ~1u;
    if (active && expand && t == 0) {
      if (p.next_pose) p.next_pose[rid] = make_float4(ex, ey, eth, ev);
    }
    if (active && t == 0 && p.final_pose && p.do_rollout && (pre & 1u))
      p.final_pose[rid] = make_float4(ex, ey, eth, ev);

    // others, second pass: commit the planned (s, v): pose on the centre line, removal,
    // Frenet table, and the pair tests against the ego's post-step rectangle
    const Box be = agent_box(scn, 0, ex, ey, ecs, esn, 0.0f, 0.0f);
    const Box bs = agent_box(scn, 0, ex, ey, ecs, esn, cfg.static_lat_safe_dist,
                             cfg.static_lon_safe_dist);
    uint32_t fl = 0, dead = 0;
    if (active) {
      for (int i = (t == 0 ? T : t); i < A; i += T) {
        const size_t gi = (size_t)i * p.n + rid;
        if (!((pre >> i) & 1u)) {
          if (expand) {
            if (p.next_pose) p.next_pose[gi] = p.pose[gi];
            if (p.reward) p.reward[gi] = 0.0f;
            if (p.last_action) p.last_action[gi] = 0.0f;
          }
          continue;
        }
        const int word = STI(L_CUR, i);
        const int cur = word & 0xff;
        const bmcts_corridor& corr = scn.corridors[cur];
        const float s_new = ST(L_NS, i);
        const int nseg = corr.n_pts - 1;
        int sg = 0;
        for (int q = 1; q < nseg; ++q)
          if (corr.s0[q] <= s_new) sg = q;
        const float ls = s_new - corr.s0[sg];
        const float nx = bm_fma(ls, corr.tx[sg], corr.px[sg]);
        const float ny = bm_fma(ls, corr.ty[sg], corr.py[sg]);
        const float nth = corr.angle[sg];
        const float nv = ST(L_NV, i);
        if (expand) {
          if (p.next_pose) p.next_pose[gi] = make_float4(nx, ny, nth, nv);
          if (p.reward) p.reward[gi] = 0.0f;
        }
        if (p.final_pose && p.do_rollout) p.final_pose[gi] = make_float4(nx, ny, nth, nv);
        if (!point_in_polygon(scn.road_x, scn.road_y, scn.n_road_pts, nx, ny)) {
          dead |= 1u << i;
          continue;
        }
        float gd, ga;
        bool gin;
        const int ncur = publish_agent(i, nx, ny, nth, false, gd, gin, ga);
        STI(L_CUR, i) = ncur | (word & ~0xff);
        ST(L_V, i) = nv;
        if (ego_alive) {
          // EvaluatorCollisionEgoAgent / EvaluatorStaticSafeDist pair (ego, i)
          const float cth = segcs[cur * BMCTS_MAX_CORRIDOR_PTS + sg];
          const float sth = segsn[cur * BMCTS_MAX_CORRIDOR_PTS + sg];
          const Box bj = agent_box(scn, i, nx, ny, cth, sth, 0.0f, 0.0f);
          if (boxes_overlap(be, bj)) fl |= BMCTS_F_COLLISION_OTHER;
          if (add_sd && boxes_overlap(bs, bj)) fl |= BMCTS_F_STATIC_SAFE_DIST;
        }
      }
    }
    if (T > 1) {
      dead = group_or<T>(dead);
      fl = group_or<T>(fl);
    }
    am &= ~dead;
    group_sync<T>();  // post-step Frenet table complete

    // --------------------------------------- mcts_observed_world_evaluation (ego view)
    float reward = 0.0f, c0 = 0.0f, c1 = 0.0f;
    uint32_t f = fl;
    if (active) {
      if (!ego_alive) {
        f |= BMCTS_F_OUT_OF_MAP;
      } else {
        // EvaluatorSafeDistDrivableArea / EvaluatorGoalReached: corner-in-polygon tests,
        // split over the lanes of the group
        const Box bd = agent_box(scn, 0, ex, ey, ecs, esn, cfg.drivable_lat_safe_dist,
                                 cfg.drivable_lon_safe_dist);
        uint32_t fc = 0;
        for (int q = t; q < 4; q += T) {
          if (!corner_in_polygon(bd, q, scn.road_x, scn.road_y, scn.n_road_pts))
            fc |= BMCTS_F_COLLISION_DRIVABLE;
          if (goal_poly && !corner_in_polygon(be, q, scn.goal.px, scn.goal.py, scn.goal.n_pts))
            fc |= kGoalFailBit;
        }
        f |= fc;
      }
    }
    if (T > 1) f = group_or<T>(f);
    if (active) {
      if (ego_alive) {
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
          const int cur = STI(L_CUR, 0) & 0xff;
          const float s_e = ST(L_PROJ + 2 * cur, 0);
          const float hwc = scn.corridors[cur].half_width;
          int jf = -1, jr = -1;
          float ds_f = kBig, ds_r = kBig;
          for (uint32_t m = am & ~1u; m; m &= m - 1u) {
            const int j = __ffs(m) - 1;
            const float dj = ST(L_PROJ + 2 * cur + 1, j);
            if (!(bm_abs(dj) <= hwc)) continue;
            const float ds = ST(L_PROJ + 2 * cur, j) - s_e;
            if (ds > 0.0f && ds < ds_f) {
              ds_f = ds;
              jf = j;
            }
            const float dr = -ds;
            if (dr > 0.0f && dr < ds_r) {
              ds_r = dr;
              jr = j;
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
    if (active) {
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
          if (p.ret_agents)
            for (int a = t; a < A; a += T) p.ret_agents[(size_t)a * p.n + rid] = (a == 0) ? R : 0.0f;
        }
      }
    }
  }
#undef ST
#undef STI
}

}  // namespace bmcts
