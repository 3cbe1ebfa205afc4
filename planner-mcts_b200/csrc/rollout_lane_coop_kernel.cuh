// rollout_lane_coop_kernel.cuh -- sm_100a rollout kernel of the cooperative state,
// "rollout-per-lane-group" mapping (see rollout_lane_kernel.cuh for the mapping itself).
//
// Replaces, for MctsStateCooperative,
//   step    = MctsStateCooperative::execute / predict / generate_next_state
//             (mcts_state/mcts_state_cooperative.cpp:45-116): every agent acts with a macro
//             action, the world is evaluated once per agent from that agent's perspective,
//             rewards are mixed with the cooperation factor
//   rollout = mamcts RandomHeuristic::calculate_heuristic_values with uniform random macro
//             actions for every agent
// (paths relative to /root/reference/bark_mcts/models/behavior/).
//
// All agents have the same role here (single-track macro action + own-perspective
// evaluation), so the T lanes of a group split them evenly.  Per-rollout state in shared
// memory, [field][agent][column]: pose, cos/sin heading, arc length on every corridor, current
// corridor, own flags, own reward, discounted return.  Every float operation mirrors
// oracle/bmcts_oracle.c.
//
// "Macro actions ahead" variant (AHEAD = true; batches too small to fill the GPU, one lane per
// agent): a macro action reads only the agent's own state and its action comes from the
// counter-based generator, so -- as for the ego of the hypothesis kernel -- a second warp computes
// EVERY agent's macro action one pass ahead of the warp that publishes, tests pairs and
// evaluates, and hands the poses over through shared memory at one named barrier per pass.  An
// agent the rollout warp removes from the map is simply not read any more.  Same arithmetic in
// the same order: bit-identical results.
#pragma once

#include "rollout_lane_kernel.cuh"

namespace bmcts {

enum { K_X = 0, K_Y, K_TH, K_V, K_CS, K_SN, K_CUR, K_FL, K_RW, K_RA, K_PROJ };  // K_PROJ + c = s

struct CoopAcq {     // the rollout a column works on, posted by the rollout warp every pass
  uint32_t seq;      // changes with every acquisition
  uint32_t rid;
  int depth;
  uint32_t flags;    // bit 0: expansion step first, bit 1: there is something to do
  uint32_t am;       // agents alive in the leaf state
  uint32_t pad_[3];
};
struct CoopMail {    // double-buffered by pass parity, like EgoMail
  uint32_t done[4];
  CoopAcq acq[2][32];
  float slot[2][32][8];  // [parity][column * T + agent]: x, y, th, v, cos, sin after the pass
};

__host__ __device__ inline LaneLayout lane_coop_layout(int A, int C, int T, int warps) {
  LaneLayout l = lane_layout(A, C, 0, T, warps);
  l.warp_bytes = sizeof(float) * (size_t)(K_PROJ + C) * A * (32 / T);
  l.total = l.off_state + l.warp_bytes * warps;
  return l;
}

template <int T, bool AHEAD>
__global__ void __launch_bounds__(AHEAD ? 64 : BMCTS_LANE_MAX_THREADS, AHEAD ? 1 : BMCTS_LANE_MIN_BLOCKS)
rollout_lane_coop_kernel(const KernelArgs p, unsigned int* __restrict__ counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NC = 32 / T;
  const int A = p.A, C = p.C;  // AHEAD: A <= T (engine.cu), one agent per lane
  const int warps = AHEAD ? 1 : blockDim.x >> 5;
  const LaneLayout lay = lane_coop_layout(A, C, T, warps);

  // ---- stage scenario prefix, macro actions, config
  {
    const uint2* src = reinterpret_cast<const uint2*>(p.scn);
    uint2* dst = reinterpret_cast<uint2*>(smem_raw);
    for (int i = threadIdx.x; i < (int)(kScnPrefixBytes / 8); i += blockDim.x) dst[i] = src[i];
    const uint32_t* s1 = reinterpret_cast<const uint32_t*>(p.scn->ego_actions);
    uint32_t* d1 = reinterpret_cast<uint32_t*>(smem_raw + lay.off_acts);
    for (int i = threadIdx.x; i < (int)(sizeof(bmcts_ego_action) * BMCTS_MAX_EGO_ACTIONS / 4);
         i += blockDim.x)
      d1[i] = s1[i];
    const uint32_t* s2 = reinterpret_cast<const uint32_t*>(p.cfg);
    uint32_t* d2 = reinterpret_cast<uint32_t*>(smem_raw + lay.off_cfg);
    for (int i = threadIdx.x; i < (int)(sizeof(bmcts_config) / 4); i += blockDim.x) d2[i] = s2[i];
    stage_reach(p.scn, p.cfg, A, reinterpret_cast<float*>(smem_raw + lay.off_rad), threadIdx.x,
                blockDim.x);
  }
  __syncthreads();  // the only block barrier: staging done
  const bmcts_scenario& scn = *reinterpret_cast<const bmcts_scenario*>(smem_raw);
  const bmcts_ego_action* acts = reinterpret_cast<const bmcts_ego_action*>(smem_raw + lay.off_acts);
  const bmcts_config& cfg = *reinterpret_cast<const bmcts_config*>(smem_raw + lay.off_cfg);
  const float* reach = reinterpret_cast<const float*>(smem_raw + lay.off_rad);
  const float* goal_bb = reach + 2 * BMCTS_MAX_AGENTS;
  const float* rects = goal_bb + 4;
  const float* goal_bb_extra = rects + 8 + 2 * BMCTS_MAX_CORRIDORS * BMCTS_MAX_CORRIDOR_PTS;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = lane % T;
  const int g = lane / T;
  CoopMail* mail = reinterpret_cast<CoopMail*>(smem_raw + smem_align16(lay.total));
  if (AHEAD && warp == 1) {
    // ------------------------------------------------------ the macro-action warp
    // lane = column * T + agent
    uint32_t seq = 0, rid = 0;
    int n = 0, depth = 1;
    bool expand = false, alive = false;
    float x = 0.0f, y = 0.0f, th = 0.0f, v = 0.0f, cs = 1.0f, sn = 0.0f;
    for (uint32_t it = 0;; ++it) {
      ahead_barrier();  // acquisitions of this pass posted, slots of this pass complete
      if (mail->done[it & 1u]) break;
      if (t < A) {
        const CoopAcq q = mail->acq[it & 1u][g];
        if (q.seq != seq) {
          seq = q.seq;
          rid = q.rid;
          depth = q.depth;
          expand = (q.flags & 1u) != 0;
          alive = (q.flags & 2u) != 0 && ((q.am >> t) & 1u) != 0;
          n = 0;
          const float4 q0 = start_pose(p, t, rid);
          x = q0.x;
          y = q0.y;
          th = q0.z;
          v = q0.w;
          bm_sincos(th, &sn, &cs);
        }
        if (alive) {
          // the corridor the publish phase finds for this pose (arg min |d|, first on ties, 0)
          float best = kBig;
          int cur = 0;
#pragma unroll 1
          for (int c = 0; c < C; ++c) {
            const Proj pr = project(scn.corridors[c], x, y);
            const float m = pr.in ? bm_abs(pr.d) : kBig;
            const bool better = m < best;
            best = better ? m : best;
            cur = better ? c : cur;
          }
          const float Dt = p.dt_table
                               ? p.dt_table[depth < BMCTS_MAX_DEPTH ? depth : BMCTS_MAX_DEPTH - 1]
                               : cfg.prediction_k;
          const size_t gi = (size_t)t * p.n + rid;
          int action;
          if (expand) {
            action = p.action[gi];
          } else {
            uint4 rr = draw(p.key0, p.key1, p.first_rollout_id + (uint64_t)rid, (uint32_t)t,
                            BMCTS_RNG_ROLLOUT_ACTION, (uint32_t)n);
            action = (int)bm_bounded(rr.x, (uint32_t)scn.n_ego_actions);
          }
          const float acc = ego_macro_step(scn, acts, cfg, action, cur, Dt, x, y, th, v, sn, cs);
          if (expand && p.last_action) p.last_action[gi] = acc;
          if (expand) expand = false; else n++;
          depth++;
          float* s = mail->slot[(it + 1u) & 1u][lane];
          s[0] = x;
          s[1] = y;
          s[2] = th;
          s[3] = v;
          s[4] = cs;
          s[5] = sn;
        }
      }
    }
    return;
  }
  uint32_t pass = 0, acq_seq = 0;  // AHEAD: pass counter (slot parity), acquisitions made
  float* col = reinterpret_cast<float*>(smem_raw + lay.off_state + lay.warp_bytes * warp) + g;
  const int fstride = A * NC;
#define FLD(f, a) col[(f) * fstride + (a) * NC]

  const uint32_t n_groups = gridDim.x * (uint32_t)warps * NC;
  const uint32_t my_group = (blockIdx.x * (uint32_t)warps + warp) * NC + g;
  const uint32_t gmask = T == 32 ? 0xffffffffu : (((1u << T) - 1u) << (lane & ~(T - 1)));

  const bool add_sd = cfg.add_safe_dist != 0;
  const int n_road = scn.n_road_pts;
  const bool want_final = p.final_pose != nullptr && p.do_rollout != 0;

  bool active = false, done = false, first = true, expand = false, fresh = false;
  uint32_t rid = 0, am = 0;
  int n = 0, depth = 1, agent_steps = 0;
  float disc = 0.0f, R = 0.0f, C0 = 0.0f, L = 0.0f;
  uint32_t f_any = 0, f_last = 0;
  uint32_t inc0 = 0, inc1 = 0, inc2 = 0, inc3 = 0;  // agents inside corridor 0..3
#define INC(c) ((c) == 0 ? inc0 : (c) == 1 ? inc1 : (c) == 2 ? inc2 : inc3)

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
        R = C0 = L = 0.0f;
        f_any = f_last = 0;
        inc0 = inc1 = inc2 = inc3 = 0;
        acq_seq++;
      }
    }
    {
      const bool all_done = __all_sync(0xffffffffu, done);
      if (AHEAD) {
        if (t == 0) {
          CoopAcq q;
          q.seq = acq_seq;
          q.rid = rid;
          q.depth = depth;
          q.flags = (expand ? 1u : 0u) | (active ? 2u : 0u);
          q.am = am;
          q.pad_[0] = q.pad_[1] = q.pad_[2] = 0u;
          mail->acq[pass & 1u][g] = q;
        }
        if (lane == 0) mail->done[pass & 1u] = all_done ? 1u : 0u;
        ahead_barrier();  // the other warp sees the acquisitions; its slots for this pass are written
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
    const uint64_t rollout_id = p.first_rollout_id + (uint64_t)rid;

    // ------------------------------------------- predict() + post-step world of own agents
    // A macro action reads only the agent's own state, so the synchronous update can be done
    // in place.  Own-perspective parts of the evaluation that do not involve another agent
    // (drivable area, goal) are done here too.
    uint32_t dead = 0, pubs = 0, in0 = 0, in1 = 0, in2 = 0, in3 = 0;
    if (active) {
#pragma unroll 1
      for (int a = t; a < A; a += T) {
        const size_t gi = (size_t)a * p.n + rid;
        float* ca = col + a * NC;
        const bool was_alive = (pre >> a) & 1u;
        float x, y, th, v, cs, sn, la = 0.0f;
        if (fresh) {
          const float4 q = start_pose(p, a, rid);
          x = q.x;
          y = q.y;
          th = q.z;
          v = q.w;
          if (want_final) p.final_pose[gi] = q;
          ca[K_RA * fstride] = 0.0f;
          ca[K_RW * fstride] = 0.0f;
          if (!was_alive) continue;
          bm_sincos(th, &sn, &cs);
        } else {
          if (!was_alive) {
            if (expand) {
              if (p.next_pose) p.next_pose[gi] = start_pose(p, a, rid);
              if (p.last_action) p.last_action[gi] = 0.0f;
            }
            continue;
          }
          if (AHEAD) {  // computed one pass ago by the macro-action warp (which also wrote
                        // last_action of an expansion step)
            const float* sl = mail->slot[pass & 1u][lane];
            x = sl[0];
            y = sl[1];
            th = sl[2];
            v = sl[3];
            cs = sl[4];
            sn = sl[5];
          } else {
            // BehaviorMPMacroActions primitive on the single-track model (spec item 5)
            int action;
            if (expand) {
              action = p.action[gi];
            } else {
              uint4 rr = draw(p.key0, p.key1, rollout_id, (uint32_t)a, BMCTS_RNG_ROLLOUT_ACTION,
                              (uint32_t)n);
              action = (int)bm_bounded(rr.x, (uint32_t)scn.n_ego_actions);
            }
            x = ca[K_X * fstride];
            y = ca[K_Y * fstride];
            th = ca[K_TH * fstride];
            v = ca[K_V * fstride];
            cs = ca[K_CS * fstride];
            sn = ca[K_SN * fstride];
            la = ego_macro_step(scn, acts, cfg, action, __float_as_int(ca[K_CUR * fstride]), Dt, x,
                                y, th, v, sn, cs);
            if (expand && p.last_action) p.last_action[gi] = la;
          }
          if (expand && p.next_pose) p.next_pose[gi] = make_float4(x, y, th, v);
          if (want_final) p.final_pose[gi] = make_float4(x, y, th, v);
          // World::RemoveInvalidAgents (spec item 6)
          if (!in_rects(rects, x, y) &&
              !point_in_polygon(scn.road_x, scn.road_y, n_road, x, y)) {
            dead |= 1u << a;
            continue;
          }
        }
        ca[K_X * fstride] = x;
        ca[K_Y * fstride] = y;
        ca[K_TH * fstride] = th;
        ca[K_V * fstride] = v;
        ca[K_CS * fstride] = cs;
        ca[K_SN * fstride] = sn;
        // the goal of THIS agent (mcts_state_cooperative.cpp:76-84: every agent is evaluated in
        // its own ObservedWorld, i.e. against its own goal definition)
        const int gk = scn.agent_goal[a];
        const bmcts_goal& gl = gk > 0 ? scn.extra_goals[gk - 1] : scn.goal;
        const float* gbb = gk > 0 ? goal_bb_extra + 4 * (gk - 1) : goal_bb;
        const bool goal_poly = gl.kind == BMCTS_GOAL_POLYGON;
        const bool goal_frenet = gl.kind == BMCTS_GOAL_FRENET_LIMITS;
        // Frenet projection on every lane corridor; current corridor = arg min |d|
        float best = kBig, goal_d = 0.0f, goal_ang = 0.0f;
        bool goal_in = false;
        int bc = 0;
        const uint32_t bit = 1u << a;
        pubs |= bit;
#pragma unroll 1
        for (int c = 0; c < C; ++c) {
          const bmcts_corridor& cc = scn.corridors[c];
          const Proj pr = project(cc, x, y);
          const float ad = bm_abs(pr.d);
          const float m = pr.in ? ad : kBig;
          ca[(K_PROJ + c) * fstride] = pr.s;
          const uint32_t inb = (pr.in && ad <= cc.half_width) ? bit : 0u;
          if (c == 0) in0 |= inb;
          if (c == 1) in1 |= inb;
          if (c == 2) in2 |= inb;
          if (c == 3) in3 |= inb;
          const bool better = m < best;
          best = better ? m : best;
          bc = better ? c : bc;
          if (goal_frenet && c == gl.corridor) {
            goal_d = pr.d;
            goal_in = pr.in;
            goal_ang = bm_norm_angle(th - cc.angle[pr.seg]);
          }
        }
        ca[K_CUR * fstride] = __int_as_float(bc);
        if (fresh) continue;
        // own-perspective evaluation that involves no other agent
        uint32_t f = 0;
        const Box bd = agent_box(scn, a, x, y, cs, sn, cfg.drivable_lat_safe_dist,
                                 cfg.drivable_lon_safe_dist);
        const Box be = agent_box(scn, a, x, y, cs, sn, 0.0f, 0.0f);
        const int n_items = goal_poly ? 8 : 4;
#pragma unroll 1
        for (int q = 0; q < n_items; ++q) {
          const bool road = q < 4;
          if (!road && (f & kGoalFailBit)) continue;  // one corner outside: goal not reached
          float cx, cy;
          box_corner(road ? bd : be, q & 3, cx, cy);
          bool in;
          if (!road && (cx < gbb[0] || cx > gbb[1] || cy < gbb[2] || cy > gbb[3]))
            in = false;  // outside the goal's bounding box (1 cm margin)
          else if (road && in_rects(rects, cx, cy))
            in = true;  // inside an inner rectangle of the road polygon
          else
            in = point_in_polygon(road ? scn.road_x : gl.px, road ? scn.road_y : gl.py,
                                  road ? n_road : gl.n_pts, cx, cy);
          if (!in) f |= road ? (uint32_t)BMCTS_F_COLLISION_DRIVABLE : kGoalFailBit;
        }
        if (goal_poly && !(f & kGoalFailBit)) f |= BMCTS_F_GOAL_REACHED;
        if (goal_frenet) {
          bool ok = goal_in && goal_d <= gl.max_lat_left && -goal_d <= gl.max_lat_right &&
                    goal_ang <= gl.max_angle_left && -goal_ang <= gl.max_angle_right &&
                    v >= gl.vel_lo && v <= gl.vel_hi;
          if (ok) f |= BMCTS_F_GOAL_REACHED;
        }
        ca[K_FL * fstride] = __uint_as_float(f & 0xffu);
      }
    }
    if (T > 1) {
      dead = group_or<T>(dead);
      pubs = group_or<T>(pubs);
      in0 = group_or<T>(in0);
      if (C > 1) in1 = group_or<T>(in1);
      if (C > 2) in2 = group_or<T>(in2);
      if (C > 3) in3 = group_or<T>(in3);
    }
    am &= ~dead;
    inc0 = (inc0 & ~pubs) | in0;
    inc1 = (inc1 & ~pubs) | in1;
    inc2 = (inc2 & ~pubs) | in2;
    inc3 = (inc3 & ~pubs) | in3;
    group_sync<T>();  // post-step world complete

    // ------------------ mcts_observed_world_evaluation from every agent's own perspective
    uint32_t f_ego = 0;
    if (stepping) {
#pragma unroll 1
      for (int a = t; a < A; a += T) {
        float* ca = col + a * NC;
        const bool alive = (am >> a) & 1u;
        uint32_t f = 0;
        if (!alive) {
          f = BMCTS_F_OUT_OF_MAP;
        } else {
          f = __float_as_uint(ca[K_FL * fstride]);
          const float x = ca[K_X * fstride], y = ca[K_Y * fstride];
          const float cs = ca[K_CS * fstride], sn = ca[K_SN * fstride];
          const Box be = agent_box(scn, a, x, y, cs, sn, 0.0f, 0.0f);
          const Box bs = agent_box(scn, a, x, y, cs, sn, cfg.static_lat_safe_dist,
                                   cfg.static_lon_safe_dist);
          // only agents within reach of this agent's (inflated) rectangle need the
          // separating-axis tests (stage_reach): first collect them, then test them
          uint32_t near = 0;
          const float ra = reach[BMCTS_MAX_AGENTS + a];
#pragma unroll 1
          for (uint32_t m = am & ~(1u << a); m; m &= m - 1u) {
            const int j = __ffs(m) - 1;
            const float thr = ra + reach[j];
            if (bm_abs(col[K_X * fstride + j * NC] - x) <= thr &&
                bm_abs(col[K_Y * fstride + j * NC] - y) <= thr)
              near |= 1u << j;
          }
#pragma unroll 1
          for (uint32_t m = near; m; m &= m - 1u) {
            const int j = __ffs(m) - 1;
            const float* cj = col + j * NC;
            const Box bj = agent_box(scn, j, cj[K_X * fstride], cj[K_Y * fstride],
                                     cj[K_CS * fstride], cj[K_SN * fstride], 0.0f, 0.0f);
            if (boxes_overlap(be, bj)) f |= BMCTS_F_COLLISION_OTHER;
            if (add_sd && boxes_overlap(bs, bj)) f |= BMCTS_F_STATIC_SAFE_DIST;
          }
          if (add_sd) {
            // EvaluatorDynamicSafeDist: nearest agents ahead of / behind in the own corridor
            const int cur = __float_as_int(ca[K_CUR * fstride]);
            const float* tab_s = col + (K_PROJ + cur) * fstride;
            const float s_e = tab_s[a * NC];
            const float v = ca[K_V * fstride];
            int jf = -1, jr = -1;
            float ds_f = kBig, ds_r = kBig;
#pragma unroll 1
            for (uint32_t m = am & INC(cur) & ~(1u << a); m; m &= m - 1u) {
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
            bool safe = true;
            if (jf >= 0) {
              float dist = ds_f - ((scn.agent_length[a] - scn.agent_rear[a]) + scn.agent_rear[jf]);
              safe = safe && safe_dist_longitudinal(FLD(K_V, jf), v, dist, -cfg.dyn_max_ego_decel,
                                                    -cfg.dyn_max_other_decel,
                                                    cfg.dyn_reaction_time_ego);
            }
            if (jr >= 0 && cfg.dyn_to_rear) {
              float dist = ds_r - ((scn.agent_length[jr] - scn.agent_rear[jr]) + scn.agent_rear[a]);
              safe = safe && safe_dist_longitudinal(v, FLD(K_V, jr), dist, -cfg.dyn_max_other_decel,
                                                    -cfg.dyn_max_ego_decel,
                                                    cfg.dyn_reaction_time_others);
            }
            if (!safe) f |= BMCTS_F_DYNAMIC_SAFE_DIST;
          }
        }
        f |= terminal_from_flags(cfg, f);
        // own reward; only agents still in the map are counted (cooperative.cpp:76-84)
        ca[K_RW * fstride] = (alive || a == 0) ? reward_from_flags(cfg, f, Dt) : 0.0f;
        if (a == 0) f_ego = f;
      }
    }
    if (T > 1) f_ego = __shfl_sync(0xffffffffu, f_ego, lane & ~(T - 1));
    group_sync<T>();  // own rewards complete

    // ------------------------- generate_next_state: reward mixing (cooperative.cpp:86-111)
    if (stepping) {
      const float r_ego = FLD(K_RW, 0);
      float sum = r_ego;
#pragma unroll 1
      for (uint32_t m = am & ~1u; m; m &= m - 1u) sum = sum + FLD(K_RW, __ffs(m) - 1);
      const uint32_t others = ~1u;
      const int n_in = __popc(am & others);
      const uint32_t missing = pre & ~am & others;
      const float cf = cfg.cooperation_factor;
      float reward0 = 0.0f;
      {
        const int n_map = n_in + __popc(missing);
        const float rest = sum - r_ego;
        const float mix = (1.0f - cf) * r_ego + cf * rest;
        reward0 = bm_div(1.0f, (float)(n_map + 1)) * mix;
      }
#pragma unroll 1
      for (int a = (t == 0 ? T : t); a < A; a += T) {
        float reward_a = 0.0f;
        if ((pre >> a) & 1u) {
          const int n_map = n_in + __popc(missing & ((2u << a) - 1u));
          const float own = ((am >> a) & 1u) ? FLD(K_RW, a) : 0.0f;
          const float rest = sum - own;
          const float mix = (1.0f - cf) * own + cf * rest;
          reward_a = bm_div(1.0f, (float)(n_map + 1)) * mix;
        }
        if (expand) {
          if (p.reward) p.reward[(size_t)a * p.n + rid] = reward_a;
        } else {
          FLD(K_RA, a) = bm_fma(disc, reward_a, FLD(K_RA, a));
        }
      }
      const float c0 = -reward0;

      // ------------------------------------------------ outputs / accumulation
      bool finished = false;
      if (expand) {
        if (t == 0) {
          if (p.reward) p.reward[rid] = reward0;
          if (p.next_alive) p.next_alive[rid] = am;
          if (p.cost) {
            p.cost[rid] = c0;
            p.cost[(size_t)p.n + rid] = 0.0f;
          }
          if (p.flags) p.flags[rid] = f_ego;
        }
        if (!p.do_rollout || (f_ego & BMCTS_F_TERMINAL) || cfg.max_rollout_steps <= 0) {
          finished = true;
          f_last = f_ego;
          f_any = f_ego;
        }
        expand = false;
      } else {
        agent_steps += __popc(pre);
        R = bm_fma(disc, reward0, R);
        C0 = bm_fma(disc, c0, C0);
        L = L + Dt;
        f_last = f_ego;
        f_any |= f_ego;
        if ((f_ego & BMCTS_F_TERMINAL) || n + 1 >= cfg.max_rollout_steps) finished = true;
        disc = disc * cfg.discount_factor;
        n++;
      }
      depth++;
      if (finished) {
        active = false;
        if (p.do_rollout) {
          if (t == 0 && p.result) {
            float4* dst = reinterpret_cast<float4*>(&p.result[rid]);
            dst[0] = make_float4(R, C0, 0.0f, L);
            dst[1] = make_float4(__uint_as_float((f_last & 0xffu) | ((f_any & 0xffu) << 8)),
                                 __int_as_float(n), __int_as_float(agent_steps),
                                 __uint_as_float(am));
          }
          if (p.ret_agents) {
#pragma unroll 1
            for (int a = t; a < A; a += T)
              p.ret_agents[(size_t)a * p.n + rid] = (a == 0) ? R : FLD(K_RA, a);
          }
        }
      }
    }
    fresh = false;
    pass++;
  }
#undef FLD
#undef INC
}

}  // namespace bmcts
