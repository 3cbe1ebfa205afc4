/*
 * bmcts_oracle.c -- scalar CPU restatement of the MCTS leaf-evaluation path of
 * bark-simulator/planner-mcts.  TEST INFRASTRUCTURE ONLY (see bmcts_oracle.h).
 * "parity unpinned" at the BARK / mamcts boundary (see bmcts_oracle.h).
 *
 * Written in the "natural" layout of the reference: one world = an array of
 * agent records, one function per reference function, plain loops.  All
 * arithmetic is FP32; elementary functions come from include/bmcts_math.h so
 * that the CUDA path can match bit for bit.  Compile with -ffp-contract=off.
 *
 * Paths cited are relative to /root/reference/bark_mcts/models/behavior/.
 */
#include "bmcts_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "bmcts_math.h"

#define MAXA BMCTS_MAX_AGENTS
#define MAXC BMCTS_MAX_CORRIDORS
#define BIG_F 1.0e30f
#define GAP_EPS 1.0e-3f

/* ------------------------------------------------------------------ RNG */

static inline void mulhilo32(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
  uint64_t p = (uint64_t)a * (uint64_t)b;
  *hi = (uint32_t)(p >> 32);
  *lo = (uint32_t)p;
}

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    mulhilo32(0xD2511F53u, c0, &hi0, &lo0);
    mulhilo32(0xCD9E8D57u, c2, &hi1, &lo1);
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n1 = lo1;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    uint32_t n3 = lo0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

static void draw(uint64_t seed, uint32_t stream, uint64_t id, uint32_t agent, uint32_t domain,
                 uint32_t index, uint32_t out[4]) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32) ^ stream};
  uint32_t ctr[4] = {(uint32_t)id, (uint32_t)(id >> 32), agent | (domain << 16), index};
  oracle_philox4x32_10(ctr, key, out);
}

void oracle_math_probe(int op, const float* a, const float* b, int n, float* out) {
  for (int i = 0; i < n; ++i) {
    float x = a[i], y = b[i], s, c, r = 0.0f;
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
}

/* ------------------------------------------- flags -> reward / cost / terminal */

float oracle_prediction_time_span(const bmcts_config* cfg, unsigned depth) {
  /* mcts_state_base.hpp:238-241: float K * std::pow(unsigned depth, float alpha) */
  return (float)((double)cfg->prediction_k * pow((double)depth, (double)cfg->prediction_alpha));
}

float oracle_reward_from_flags(const bmcts_config* cfg, uint32_t f, float dt) {
  /* mcts_state_base.hpp:107-118 */
  int coll = (f & BMCTS_F_COLLISION_OTHER) != 0;
  int stat = (f & BMCTS_F_STATIC_SAFE_DIST) != 0;
  int dyn = (f & BMCTS_F_DYNAMIC_SAFE_DIST) != 0;
  int driv = (f & BMCTS_F_COLLISION_DRIVABLE) != 0;
  int goal = (f & BMCTS_F_GOAL_REACHED) != 0;
  int oom = (f & BMCTS_F_OUT_OF_MAP) != 0;
  int sd_pred = cfg->static_safe_dist_is_terminal ? dyn : (dyn || stat);
  float safe_dist_reward = ((float)sd_pred * cfg->safe_dist_violated_reward) * dt;
  int coll_pred = coll || (cfg->static_safe_dist_is_terminal ? stat : 0);
  float collision_reward =
      (float)coll_pred * cfg->collision_reward + (float)(driv || oom) * cfg->out_of_drivable_reward;
  float r = bm_max(collision_reward, cfg->collision_reward);
  r = r + (cfg->add_safe_dist ? safe_dist_reward : 0.0f);
  r = r + (float)goal * cfg->goal_reward;
  r = r + cfg->step_reward;
  return r;
}

int oracle_costs_from_flags(const bmcts_config* cfg, uint32_t f, float dt, float cost[2]) {
  /* mcts_state_base.hpp:120-142 */
  int coll = (f & BMCTS_F_COLLISION_OTHER) != 0;
  int stat = (f & BMCTS_F_STATIC_SAFE_DIST) != 0;
  int dyn = (f & BMCTS_F_DYNAMIC_SAFE_DIST) != 0;
  int driv = (f & BMCTS_F_COLLISION_DRIVABLE) != 0;
  int goal = (f & BMCTS_F_GOAL_REACHED) != 0;
  int oom = (f & BMCTS_F_OUT_OF_MAP) != 0;
  int sd_pred = cfg->static_safe_dist_is_terminal ? dyn : (dyn || stat);
  float safe_dist_cost = ((float)sd_pred * cfg->safe_dist_violated_cost) * dt;
  int coll_pred = coll || (cfg->static_safe_dist_is_terminal ? stat : 0) ||
                  (cfg->dynamic_safe_dist_is_terminal ? dyn : 0);
  float total = (float)coll_pred * cfg->collision_cost;
  total = total + (float)(driv || oom) * cfg->out_of_drivable_cost;
  total = total +
          ((cfg->add_safe_dist && !cfg->split_safe_dist_collision) ? safe_dist_cost : 0.0f);
  total = total + (float)goal * cfg->goal_cost;
  cost[0] = 0.0f;
  cost[1] = 0.0f;
  if (cfg->split_safe_dist_collision) {
    cost[0] = cfg->chance_costs ? bm_min(safe_dist_cost, 1.0f) : safe_dist_cost;
    cost[1] = bm_min(total, cfg->collision_cost) * dt;
    return 2;
  }
  cost[0] = cfg->chance_costs ? bm_min(total, 1.0f) : total;
  return 1;
}

uint32_t oracle_terminal_from_flags(const bmcts_config* cfg, uint32_t f) {
  /* mcts_state_base.hpp:273-277 */
  int t = (f & (BMCTS_F_COLLISION_OTHER | BMCTS_F_OUT_OF_MAP | BMCTS_F_GOAL_REACHED)) != 0;
  t = t || ((f & BMCTS_F_COLLISION_DRIVABLE) && cfg->out_of_drivable_is_terminal);
  t = t || ((f & BMCTS_F_STATIC_SAFE_DIST) && cfg->static_safe_dist_is_terminal);
  t = t || ((f & BMCTS_F_DYNAMIC_SAFE_DIST) && cfg->dynamic_safe_dist_is_terminal);
  return t ? BMCTS_F_TERMINAL : 0u;
}

/* ------------------------------------------------------------- geometry */

void oracle_project(const bmcts_corridor* c, float x, float y, float* s_out, float* d_out,
                    int* seg_out, int* in_range_out) {
  /* nearest point on the polyline; first minimal segment wins on ties (spec item 4) */
  int nseg = c->n_pts - 1;
  float best_d2 = BIG_F, best_u = 0.0f, best_raw = 0.0f, best_cross = 0.0f;
  int best = 0;
  for (int k = 0; k < nseg; ++k) {
    float rx = x - c->px[k];
    float ry = y - c->py[k];
    float u = bm_fma(rx, c->tx[k], ry * c->ty[k]);
    float uc = bm_clamp(u, 0.0f, c->seg_len[k]);
    float ex = bm_fma(-uc, c->tx[k], rx);
    float ey = bm_fma(-uc, c->ty[k], ry);
    float d2 = bm_fma(ex, ex, ey * ey);
    if (d2 < best_d2) {
      best_d2 = d2;
      best = k;
      best_u = uc;
      best_raw = u;
      best_cross = bm_fma(c->tx[k], ry, -(c->ty[k] * rx));
    }
  }
  *s_out = c->s0[best] + best_u;
  *d_out = best_cross;
  *seg_out = best;
  int in_range = 1;
  if (best == 0 && best_raw < 0.0f) in_range = 0;
  if (best == nseg - 1 && best_raw > c->seg_len[best]) in_range = 0;
  *in_range_out = in_range;
}

static int point_in_polygon(const float* px, const float* py, int n, float x, float y) {
  /* even-odd rule, division free */
  int inside = 0;
  for (int i = 0, j = n - 1; i < n; j = i++) {
    float yi = py[i], yj = py[j];
    if ((yi > y) != (yj > y)) {
      float lhs = (x - px[i]) * (yj - yi);
      float rhs = (px[j] - px[i]) * (y - yi);
      int cross = (yj > yi) ? (lhs < rhs) : (lhs > rhs);
      inside ^= cross;
    }
  }
  return inside;
}

typedef struct {
  float cx, cy, c, s, hl, hw;
} box_t;

/* rectangle of an agent, inflated by lon at both ends and lat at both sides */
static box_t agent_box(const bmcts_scenario* scn, int a, float x, float y, float c, float s,
                       float lat, float lon) {
  float front = (scn->agent_length[a] - scn->agent_rear[a]) + lon;
  float back = scn->agent_rear[a] + lon;
  box_t b;
  b.hl = 0.5f * (front + back);
  float off = 0.5f * (front - back);
  b.cx = bm_fma(off, c, x);
  b.cy = bm_fma(off, s, y);
  b.c = c;
  b.s = s;
  b.hw = bm_fma(0.5f, scn->agent_width[a], lat);
  return b;
}

/* separating-axis test of two rectangles; touching counts as separated */
static int boxes_overlap(const box_t* A, const box_t* B) {
  float tx = B->cx - A->cx;
  float ty = B->cy - A->cy;
  float cd = bm_fma(A->c, B->c, A->s * B->s);
  float sd = bm_fma(A->c, B->s, -(A->s * B->c));
  float acd = bm_abs(cd), asd = bm_abs(sd);
  float p0 = bm_abs(bm_fma(tx, A->c, ty * A->s));
  if (p0 >= A->hl + bm_fma(B->hl, acd, B->hw * asd)) return 0;
  float p1 = bm_abs(bm_fma(ty, A->c, -(tx * A->s)));
  if (p1 >= A->hw + bm_fma(B->hl, asd, B->hw * acd)) return 0;
  float p2 = bm_abs(bm_fma(tx, B->c, ty * B->s));
  if (p2 >= B->hl + bm_fma(A->hl, acd, A->hw * asd)) return 0;
  float p3 = bm_abs(bm_fma(ty, B->c, -(tx * B->s)));
  if (p3 >= B->hw + bm_fma(A->hl, asd, A->hw * acd)) return 0;
  return 1;
}

static void box_corner(const box_t* b, int k, float* x, float* y) {
  float lx = (k == 0 || k == 1) ? b->hl : -b->hl;
  float ly = (k == 0 || k == 3) ? b->hw : -b->hw;
  *x = b->cx + (lx * b->c - ly * b->s);
  *y = b->cy + (lx * b->s + ly * b->c);
}

static int box_within_polygon(const box_t* b, const float* px, const float* py, int n) {
  for (int k = 0; k < 4; ++k) {
    float x, y;
    box_corner(b, k, &x, &y);
    if (!point_in_polygon(px, py, n, x, y)) return 0;
  }
  return 1;
}

/* ---------------------------------------------------- world snapshot */

typedef struct {
  int alive;
  float x, y, th, v;
  float c, s; /* cos / sin heading */
  float ps[MAXC], pd[MAXC];
  int pseg[MAXC], pin[MAXC];
  int cur; /* current lane corridor (spec item 4) */
} agent_t;

typedef struct {
  int n;
  agent_t a[MAXA];
} world_t;

static void publish(const bmcts_scenario* scn, world_t* w) {
  for (int i = 0; i < w->n; ++i) {
    agent_t* a = &w->a[i];
    if (!a->alive) continue;
    bm_sincos(a->th, &a->s, &a->c);
    float best = BIG_F;
    int cur = 0;
    for (int c = 0; c < scn->n_corridors; ++c) {
      oracle_project(&scn->corridors[c], a->x, a->y, &a->ps[c], &a->pd[c], &a->pseg[c],
                     &a->pin[c]);
      float m = a->pin[c] ? bm_abs(a->pd[c]) : BIG_F;
      if (m < best) {
        best = m;
        cur = c;
      }
    }
    a->cur = cur;
  }
}

static void load_world(const bmcts_scenario* scn, const float* pose, uint32_t alive, world_t* w) {
  w->n = scn->n_agents;
  for (int i = 0; i < w->n; ++i) {
    agent_t* a = &w->a[i];
    memset(a, 0, sizeof(*a));
    a->alive = (alive >> i) & 1u;
    a->x = pose[4 * i + 0];
    a->y = pose[4 * i + 1];
    a->th = pose[4 * i + 2];
    a->v = pose[4 * i + 3];
  }
  publish(scn, w);
}

static int in_corridor(const bmcts_scenario* scn, const agent_t* a, int c) {
  return a->pin[c] && bm_abs(a->pd[c]) <= scn->corridors[c].half_width;
}

/* ------------------------------------------------------- behaviour models */

typedef struct {
  float T, s0, a_max, v0, b, cool;
} idm_params_t;

static void sample_idm_params(const bmcts_hypothesis* h, const uint32_t r0[4],
                              const uint32_t r1[4], idm_params_t* p) {
  /* BehaviorIDMStochastic::SampleParameters order (spec item 6) */
  float u[6] = {bm_u01(r0[0]), bm_u01(r0[1]), bm_u01(r0[2]),
                bm_u01(r0[3]), bm_u01(r1[0]), bm_u01(r1[1])};
  float v[6];
  for (int m = 0; m < 6; ++m) v[m] = bm_fma(u[m], h->hi[m] - h->lo[m], h->lo[m]);
  p->T = v[0];
  p->s0 = v[1];
  p->a_max = v[2];
  p->v0 = v[3];
  p->b = v[4];
  p->cool = v[5];
}

/* BARK BaseIDM::CalcRelativeValues: nearest alive agent ahead in corridor c */
static int find_leader(const bmcts_scenario* scn, const world_t* w, int i, int c, float* ds_out) {
  float s_i = w->a[i].ps[c];
  float best = BIG_F;
  int lead = -1;
  for (int j = 0; j < w->n; ++j) {
    if (j == i || !w->a[j].alive) continue;
    if (!in_corridor(scn, &w->a[j], c)) continue;
    float ds = w->a[j].ps[c] - s_i;
    if (ds > 0.0f && ds < best) {
      best = ds;
      lead = j;
    }
  }
  *ds_out = best;
  return lead;
}

/* IDM acceleration (Treiber et al. 2000), BARK BaseIDM::GetTotalAcc without
 * the coolness blend (coolness must be 0, bmcts_scenario_finalize rejects
 * anything else); clamped to the hypothesis' acceleration bounds */
static float idm_acc(const bmcts_hypothesis* h, const idm_params_t* p, float inv_v0,
                     float inv_2sab, int has_leader, float v, float v_l, float gap) {
  float free_term = 1.0f - bm_powi(v * inv_v0, h->exponent);
  float acc;
  if (has_leader) {
    float dv = v - v_l;
    float s_star = p->s0 + bm_max(0.0f, bm_fma(v, p->T, (v * dv) * inv_2sab));
    float q = bm_div(s_star, bm_max(gap, GAP_EPS));
    acc = p->a_max * (free_term - q * q);
  } else {
    acc = p->a_max * free_term;
  }
  return bm_clamp(acc, h->acc_lower_bound, h->acc_upper_bound);
}

static void idm_constants(const idm_params_t* p, float* inv_v0, float* inv_2sab) {
  *inv_v0 = bm_div(1.0f, bm_max(p->v0, 1.0e-3f));
  *inv_2sab = bm_div(1.0f, 2.0f * bm_sqrt(bm_max(p->a_max * p->b, 1.0e-6f)));
}

/* BehaviorIDMStochastic::Plan for one other agent (spec items 5-7) */
static void idm_agent_step(const bmcts_config* cfg, const bmcts_scenario* scn, const world_t* w,
                           int i, const bmcts_hypothesis* h, const idm_params_t* p, float Dt,
                           agent_t* out, float* last_action) {
  const agent_t* a = &w->a[i];
  int c = a->cur;
  const bmcts_corridor* corr = &scn->corridors[c];
  float ds_lead;
  int lead = find_leader(scn, w, i, c, &ds_lead);
  int has_leader = lead >= 0;
  float gap = 0.0f, v_l = 0.0f;
  if (has_leader) {
    gap = ds_lead - ((scn->agent_length[i] - scn->agent_rear[i]) + scn->agent_rear[lead]);
    v_l = w->a[lead].v;
  }
  float inv_v0, inv_2sab;
  idm_constants(p, &inv_v0, &inv_2sab);
  int nsub = cfg->num_substeps;
  float dt = bm_div(Dt, (float)nsub);
  float half_dt2 = (0.5f * dt) * dt;
  float v = a->v, trav = 0.0f, acc = 0.0f;
  for (int k = 0; k < nsub; ++k) {
    acc = idm_acc(h, p, inv_v0, inv_2sab, has_leader, v, v_l, gap);
    float ds = bm_max(0.0f, bm_fma(acc, half_dt2, v * dt));
    v = bm_clamp(bm_fma(acc, dt, v), h->min_velocity, h->max_velocity);
    gap = gap + bm_fma(v_l, dt, -ds);
    trav = trav + ds;
  }
  float s_new = bm_min(a->ps[c] + trav, corr->length);
  int nseg = corr->n_pts - 1;
  int k = 0;
  for (int q = 1; q < nseg; ++q)
    if (corr->s0[q] <= s_new) k = q;
  float ls = s_new - corr->s0[k];
  out->x = bm_fma(ls, corr->tx[k], corr->px[k]);
  out->y = bm_fma(ls, corr->ty[k], corr->py[k]);
  out->th = corr->angle[k];
  out->v = v;
  *last_action = acc;
}

int oracle_ego_substeps(const bmcts_config* cfg, float Dt) {
  if (!(cfg->ego_integration_dt > 0.0f)) return cfg->num_substeps;
  float k = ceilf(bm_div(Dt, cfg->ego_integration_dt) - 1.0e-3f);
  k = bm_clamp(k, 1.0f, 1024.0f);
  return (int)k;
}

/* BehaviorMPMacroActions primitive on the single-track model (spec item 8) */
static void macro_agent_step(const bmcts_config* cfg, const bmcts_scenario* scn, const world_t* w,
                             int i, int action, float Dt, agent_t* out, float* last_action) {
  const agent_t* a = &w->a[i];
  const bmcts_ego_action* act = &scn->ego_actions[action];
  float acc = bm_clamp(act->acceleration, cfg->lon_acc_min, cfg->lon_acc_max);
  int tgt = a->cur;
  if (act->type == BMCTS_ACT_CHANGE_LEFT && scn->corridors[tgt].left >= 0)
    tgt = scn->corridors[tgt].left;
  else if (act->type == BMCTS_ACT_CHANGE_RIGHT && scn->corridors[tgt].right >= 0)
    tgt = scn->corridors[tgt].right;
  const bmcts_corridor* corr = &scn->corridors[tgt];
  /* BehaviorMotionPrimitives::IntegrationTimeDelta: ceil(Dt / delta) Euler steps of equal
   * length (1e-3 of slack: 0.5 / float(0.02) is 25.0000006) */
  int nsub = oracle_ego_substeps(cfg, Dt);
  float dt = bm_div(Dt, (float)nsub);
  float inv_wb = bm_div(1.0f, cfg->wheel_base);
  float x = a->x, y = a->y, th = a->th, v = a->v;
  for (int k = 0; k < nsub; ++k) {
    float sn, cs;
    bm_sincos(th, &sn, &cs);
    float fx = bm_fma(cfg->wheel_base, cs, x);
    float fy = bm_fma(cfg->wheel_base, sn, y);
    float ps, pd;
    int seg, in;
    oracle_project(corr, fx, fy, &ps, &pd, &seg, &in);
    float th_err = bm_norm_angle(corr->angle[seg] - th);
    float delta = th_err + bm_atan2(-(cfg->crosstrack_gain * pd), v);
    delta = bm_clamp(delta, -cfg->delta_max, cfg->delta_max);
    float tn = bm_tan_small(delta);
    if (cfg->limit_steering_rate) {
      /* BARK CalculateSteeringAngle(limit_steering): lateral acceleration v^2 tan(delta) / L
       * kept inside [LatAccMin, LatAccMax], i.e. tan(delta) inside [LatAccMin, LatAccMax] L / v^2 */
      float r = bm_div(cfg->wheel_base, bm_max(v * v, 1.0e-6f));
      tn = bm_clamp(tn, cfg->lat_acc_min * r, cfg->lat_acc_max * r);
    }
    float dv = dt * v;
    x = bm_fma(dv, cs, x);
    y = bm_fma(dv, sn, y);
    th = bm_fma(dv, tn * inv_wb, th);
    v = bm_clamp(bm_fma(dt, acc, v), 0.0f, cfg->ego_max_velocity);
  }
  out->x = x;
  out->y = y;
  out->th = bm_norm_angle(th);
  out->v = v;
  *last_action = acc;
}

/* ------------------------------------------------------------ evaluation */

/* longitudinal safe distance with reaction time (Rizaldi, Immler, Althoff 2016;
 * BARK SafeDistanceLabelFunction::CheckSafeDistanceLongitudinal) */
static int safe_dist_longitudinal(float v_f, float v_r, float dist, float a_r, float a_f,
                                  float delta) {
  if (dist < 0.0f) return 0;
  float t_stop_f = bm_div(-v_f, a_f);
  float v_f_star = (delta <= t_stop_f) ? bm_fma(a_f, delta, v_f) : 0.0f;
  float t_stop_f_star = bm_div(-v_f_star, a_f); /* the front vehicle's own stopping time */
  float t_stop_r = bm_div(-v_r, a_r);
  float sd0 = v_r * delta - bm_div(v_r * v_r, 2.0f * a_r);
  float sd1 = sd0 + bm_div(v_f * v_f, 2.0f * a_f);
  float front_react = v_f * delta + (0.5f * a_f) * (delta * delta);
  float sd3 = sd0 - front_react;
  if (dist > sd0 || (delta <= t_stop_f && dist > sd3)) return 1;
  if (delta <= t_stop_f && a_f > a_r && v_f_star < v_r && t_stop_r < t_stop_f_star) {
    float rel = v_f_star - v_r;
    float sd2 = bm_div(rel * rel, 2.0f * (a_f - a_r)) - front_react + v_r * delta;
    return dist > sd2;
  }
  return dist > sd1;
}

/* mcts_observed_world_evaluation (mcts_state_base.hpp:248-279) from the
 * perspective of agent e of the (published) post-step world */
static uint32_t evaluate(const bmcts_config* cfg, const bmcts_scenario* scn, const world_t* w,
                         int e) {
  uint32_t f = 0;
  const agent_t* a = &w->a[e];
  if (!a->alive) {
    f |= BMCTS_F_OUT_OF_MAP; /* :270-272 */
  } else {
    /* :257-258 EvaluatorSafeDistDrivableArea: inflated shape within road corridor */
    box_t bd = agent_box(scn, e, a->x, a->y, a->c, a->s, cfg->drivable_lat_safe_dist,
                         cfg->drivable_lon_safe_dist);
    if (!box_within_polygon(&bd, scn->road_x, scn->road_y, scn->n_road_pts))
      f |= BMCTS_F_COLLISION_DRIVABLE;
    /* :259-260 EvaluatorCollisionEgoAgent */
    box_t be = agent_box(scn, e, a->x, a->y, a->c, a->s, 0.0f, 0.0f);
    for (int j = 0; j < w->n; ++j) {
      if (j == e || !w->a[j].alive) continue;
      box_t bj = agent_box(scn, j, w->a[j].x, w->a[j].y, w->a[j].c, w->a[j].s, 0.0f, 0.0f);
      if (boxes_overlap(&be, &bj)) f |= BMCTS_F_COLLISION_OTHER;
    }
    /* :261-262 EvaluatorGoalReached against the goal of the agent whose perspective this is
     * (mcts_state_cooperative.cpp:76-84 builds ObservedWorld(predicted_world, agent)) */
    const bmcts_goal* g = scn->agent_goal[e] > 0 ? &scn->extra_goals[scn->agent_goal[e] - 1] : &scn->goal;
    if (g->kind == BMCTS_GOAL_POLYGON) {
      if (box_within_polygon(&be, g->px, g->py, g->n_pts)) f |= BMCTS_F_GOAL_REACHED;
    } else if (g->kind == BMCTS_GOAL_FRENET_LIMITS) {
      int c = g->corridor;
      float d = a->pd[c];
      float ang = bm_norm_angle(a->th - scn->corridors[c].angle[a->pseg[c]]);
      int ok = a->pin[c] && d <= g->max_lat_left && -d <= g->max_lat_right &&
               ang <= g->max_angle_left && -ang <= g->max_angle_right && a->v >= g->vel_lo &&
               a->v <= g->vel_hi;
      if (ok) f |= BMCTS_F_GOAL_REACHED;
    }
    if (cfg->add_safe_dist) { /* :265-268 */
      /* EvaluatorStaticSafeDist: inflated ego shape collides with another agent */
      box_t bs = agent_box(scn, e, a->x, a->y, a->c, a->s, cfg->static_lat_safe_dist,
                           cfg->static_lon_safe_dist);
      for (int j = 0; j < w->n; ++j) {
        if (j == e || !w->a[j].alive) continue;
        box_t bj = agent_box(scn, j, w->a[j].x, w->a[j].y, w->a[j].c, w->a[j].s, 0.0f, 0.0f);
        if (boxes_overlap(&bs, &bj)) f |= BMCTS_F_STATIC_SAFE_DIST;
      }
      /* EvaluatorDynamicSafeDist: longitudinal safe distance to the nearest agents
       * in front of and behind the ego inside its lane corridor */
      int c = a->cur;
      float s_e = a->ps[c];
      float best_f = BIG_F, best_r = BIG_F;
      int jf = -1, jr = -1;
      for (int j = 0; j < w->n; ++j) {
        if (j == e || !w->a[j].alive) continue;
        if (!in_corridor(scn, &w->a[j], c)) continue;
        float ds = w->a[j].ps[c] - s_e;
        if (ds > 0.0f && ds < best_f) {
          best_f = ds;
          jf = j;
        }
        float dr = -ds;
        if (dr > 0.0f && dr < best_r) {
          best_r = dr;
          jr = j;
        }
      }
      int safe = 1;
      if (jf >= 0) {
        float dist = best_f - ((scn->agent_length[e] - scn->agent_rear[e]) + scn->agent_rear[jf]);
        safe = safe && safe_dist_longitudinal(w->a[jf].v, a->v, dist, -cfg->dyn_max_ego_decel,
                                              -cfg->dyn_max_other_decel,
                                              cfg->dyn_reaction_time_ego);
      }
      if (jr >= 0 && cfg->dyn_to_rear) {
        float dist = best_r - ((scn->agent_length[jr] - scn->agent_rear[jr]) + scn->agent_rear[e]);
        safe = safe && safe_dist_longitudinal(a->v, w->a[jr].v, dist, -cfg->dyn_max_other_decel,
                                              -cfg->dyn_max_ego_decel,
                                              cfg->dyn_reaction_time_others);
      }
      if (!safe) f |= BMCTS_F_DYNAMIC_SAFE_DIST;
    }
  }
  f |= oracle_terminal_from_flags(cfg, f);
  return f;
}

/* ------------------------------------------------------------- one step */

typedef struct {
  uint64_t seed;
  uint32_t stream;
  uint32_t domain;
  uint64_t id[MAXA];   /* per-agent Philox id */
  uint32_t index_base; /* ROLLOUT_PARAMS: 2*step, TREE_PARAMS: 0 */
} draw_spec_t;

/* execute(): predict -> evaluate -> rewards / costs.  w is replaced by the
 * published post-step world. */
static uint32_t world_step(const bmcts_config* cfg, const bmcts_scenario* scn, world_t* w,
                           const uint8_t* hyp_id, const uint8_t* action, const draw_spec_t* ds,
                           unsigned depth, float* reward, float* cost, float* last_action,
                           float* dt_out) {
  int A = w->n;
  int coop = cfg->state_kind == BMCTS_STATE_COOPERATIVE;
  float Dt = oracle_prediction_time_span(cfg, depth);
  *dt_out = Dt;
  world_t nw = *w;
  /* predict(): every agent plans against the pre-step snapshot (spec item 3) */
  for (int i = 0; i < A; ++i) {
    if (last_action) last_action[i] = 0.0f;
    if (!w->a[i].alive) continue;
    float la = 0.0f;
    if (i == 0 || coop) {
      macro_agent_step(cfg, scn, w, i, action[i], Dt, &nw.a[i], &la);
    } else {
      const bmcts_hypothesis* h = &scn->hypotheses[hyp_id[i]];
      uint32_t r0[4], r1[4];
      draw(ds->seed, ds->stream, ds->id[i], (uint32_t)i, ds->domain, ds->index_base + 0, r0);
      draw(ds->seed, ds->stream, ds->id[i], (uint32_t)i, ds->domain, ds->index_base + 1, r1);
      idm_params_t p;
      sample_idm_params(h, r0, r1, &p);
      idm_agent_step(cfg, scn, w, i, h, &p, Dt, &nw.a[i], &la);
    }
    if (last_action) last_action[i] = la;
  }
  /* World::RemoveInvalidAgents: reference point outside the road corridor (spec item 9) */
  for (int i = 0; i < A; ++i) {
    if (!nw.a[i].alive) continue;
    if (!point_in_polygon(scn->road_x, scn->road_y, scn->n_road_pts, nw.a[i].x, nw.a[i].y))
      nw.a[i].alive = 0;
  }
  publish(scn, &nw);
  /* evaluation + generate_next_state */
  uint32_t f_ego = evaluate(cfg, scn, &nw, 0);
  for (int i = 0; i < A; ++i) reward[i] = 0.0f;
  if (!coop) {
    /* mcts_state_hypothesis.hpp:125-140, mcts_state_risk_constraint.cpp:67-81 */
    reward[0] = oracle_reward_from_flags(cfg, f_ego, Dt);
    oracle_costs_from_flags(cfg, f_ego, Dt, cost);
  } else {
    /* mcts_state_cooperative.cpp:70-116 */
    float r_own[MAXA];
    int counted[MAXA];
    float r_ego = oracle_reward_from_flags(cfg, f_ego, Dt);
    float sum = r_ego;
    int n_in = 0;
    for (int j = 1; j < A; ++j) {
      r_own[j] = 0.0f;
      counted[j] = 0;
      /* :76 iterates predicted_world->GetOtherAgents(): agents still in the map */
      if (!nw.a[j].alive) continue;
      uint32_t fj = evaluate(cfg, scn, &nw, j);
      r_own[j] = oracle_reward_from_flags(cfg, fj, Dt);
      counted[j] = 1;
      n_in++;
    }
    for (int j = 1; j < A; ++j)
      if (counted[j]) sum = sum + r_own[j];
    float cf = cfg->cooperation_factor;
    /* :92-102: other_agents_rewards[agent_idx] inserts a zero entry for an agent of
     * THIS state that left the map, so the divisor grows as the loop proceeds */
    int n_map = n_in;
    for (int j = 1; j < A; ++j) {
      if (!w->a[j].alive) continue; /* not in get_other_agent_idx() of this state */
      if (!counted[j]) n_map++;
      float rest = sum - r_own[j];
      float mix = (1.0f - cf) * r_own[j] + cf * rest;
      reward[j] = bm_div(1.0f, (float)(n_map + 1)) * mix;
    }
    float rest = sum - r_ego;
    float mix = (1.0f - cf) * r_ego + cf * rest;
    reward[0] = bm_div(1.0f, (float)(n_map + 1)) * mix; /* :105-108 */
    cost[0] = -reward[0];                               /* :110-111 */
    cost[1] = 0.0f;
  }
  *w = nw;
  return f_ego;
}

static uint32_t alive_mask(const world_t* w) {
  uint32_t m = 0;
  for (int i = 0; i < w->n; ++i)
    if (w->a[i].alive) m |= 1u << i;
  return m;
}

static void store_poses(const world_t* w, float* pose) {
  for (int i = 0; i < w->n; ++i) {
    pose[4 * i + 0] = w->a[i].x;
    pose[4 * i + 1] = w->a[i].y;
    pose[4 * i + 2] = w->a[i].th;
    pose[4 * i + 3] = w->a[i].v;
  }
}

static int check_args(const bmcts_config* cfg, const bmcts_scenario* scn) {
  if (!cfg || !scn || !scn->finalized) return BMCTS_ERR_INVALID_ARG;
  if (scn->n_agents < 1 || scn->n_agents > MAXA) return BMCTS_ERR_INVALID_ARG;
  return BMCTS_OK;
}

int oracle_step(const bmcts_config* cfg, const bmcts_scenario* scn, const float* pose,
                uint32_t alive, const uint8_t* hyp_id, unsigned depth, const uint8_t* action,
                const uint64_t* draw_id, uint64_t seed, uint32_t stream, float* next_pose,
                uint32_t* next_alive, float* reward, float* cost, uint32_t* flags,
                float* last_action) {
  int st = check_args(cfg, scn);
  if (st) return st;
  world_t w;
  load_world(scn, pose, alive, &w);
  draw_spec_t ds;
  ds.seed = seed;
  ds.stream = stream;
  ds.domain = BMCTS_RNG_TREE_PARAMS;
  ds.index_base = 0;
  for (int i = 0; i < w.n; ++i) ds.id[i] = draw_id ? draw_id[i] : 0;
  float dt;
  float c2[2] = {0.0f, 0.0f};
  float rw[MAXA];
  uint32_t f = world_step(cfg, scn, &w, hyp_id, action, &ds, depth, rw, c2, last_action, &dt);
  if (next_pose) store_poses(&w, next_pose);
  if (next_alive) *next_alive = alive_mask(&w);
  if (reward) memcpy(reward, rw, sizeof(float) * (size_t)w.n);
  if (cost) {
    cost[0] = c2[0];
    cost[1] = c2[1];
  }
  if (flags) *flags = f;
  return BMCTS_OK;
}

/* rollout from an already loaded world */
static void rollout_world(const bmcts_config* cfg, const bmcts_scenario* scn, world_t* w,
                          const uint8_t* hyp_id, unsigned depth, uint64_t seed, uint32_t stream,
                          uint64_t rollout_id, bmcts_rollout_result* res, float* ret_agents,
                          float* final_pose) {
  int A = w->n;
  int coop = cfg->state_kind == BMCTS_STATE_COOPERATIVE;
  /* mamcts RandomHeuristic: the first rollout reward is already discounted once
   * (modified_discount_factor starts at DISCOUNT_FACTOR) */
  float disc = cfg->discount_factor;
  float R = 0.0f, C0 = 0.0f, C1 = 0.0f, L = 0.0f;
  float Ra[MAXA];
  for (int i = 0; i < A; ++i) Ra[i] = 0.0f;
  uint32_t f_last = 0, f_any = 0;
  int n = 0, agent_steps = 0;
  draw_spec_t ds;
  ds.seed = seed;
  ds.stream = stream;
  ds.domain = BMCTS_RNG_ROLLOUT_PARAMS;
  for (int i = 0; i < A; ++i) ds.id[i] = rollout_id;
  while (n < cfg->max_rollout_steps) {
    uint8_t action[MAXA];
    for (int i = 0; i < A; ++i) {
      action[i] = 0;
      if (!(i == 0 || coop) || !w->a[i].alive) continue;
      uint32_t r[4];
      draw(seed, stream, rollout_id, (uint32_t)i, BMCTS_RNG_ROLLOUT_ACTION, (uint32_t)n, r);
      action[i] = (uint8_t)bm_bounded(r[0], (uint32_t)scn->n_ego_actions);
    }
    for (int i = 0; i < A; ++i) agent_steps += w->a[i].alive;
    ds.index_base = 2u * (uint32_t)n;
    float rw[MAXA], c2[2] = {0.0f, 0.0f}, dt;
    uint32_t f = world_step(cfg, scn, w, hyp_id, action, &ds, depth, rw, c2, NULL, &dt);
    R = bm_fma(disc, rw[0], R);
    C0 = bm_fma(disc, c2[0], C0);
    C1 = bm_fma(disc, c2[1], C1);
    for (int i = 0; i < A; ++i) Ra[i] = bm_fma(disc, rw[i], Ra[i]);
    L = L + dt;
    disc = disc * cfg->discount_factor;
    depth++;
    n++;
    f_last = f;
    f_any |= f;
    if (f & BMCTS_F_TERMINAL) break;
  }
  res->ret_ego = R;
  res->cost0 = C0;
  res->cost1 = C1;
  res->step_len = L;
  res->flags = (f_last & 0xffu) | ((f_any & 0xffu) << 8);
  res->n_steps = n;
  res->agent_steps = agent_steps;
  res->final_alive = alive_mask(w);
  if (ret_agents) memcpy(ret_agents, Ra, sizeof(float) * (size_t)A);
  if (final_pose) store_poses(w, final_pose);
}

int oracle_rollout(const bmcts_config* cfg, const bmcts_scenario* scn, const float* pose,
                   uint32_t alive, const uint8_t* hyp_id, unsigned depth, uint64_t seed,
                   uint32_t stream, uint64_t rollout_id, bmcts_rollout_result* res,
                   float* ret_agents, float* final_pose) {
  int st = check_args(cfg, scn);
  if (st) return st;
  world_t w;
  load_world(scn, pose, alive, &w);
  rollout_world(cfg, scn, &w, hyp_id, depth, seed, stream, rollout_id, res, ret_agents,
                final_pose);
  return BMCTS_OK;
}

/* ------------------------------------------------------------- batches */
/* batch arrays are agent-major SoA (bmcts_c.h): element (a, k) of n is at a*n + k */

static void gather_f4(const float* src, int A, int n, int k, float* dst) {
  for (int a = 0; a < A; ++a) memcpy(dst + 4 * a, src + ((size_t)a * n + k) * 4, 4 * sizeof(float));
}
static void scatter_f4(float* dst, int A, int n, int k, const float* src) {
  for (int a = 0; a < A; ++a) memcpy(dst + ((size_t)a * n + k) * 4, src + 4 * a, 4 * sizeof(float));
}
static void gather_u8(const uint8_t* src, int A, int n, int k, uint8_t* dst) {
  for (int a = 0; a < A; ++a) dst[a] = src ? src[(size_t)a * n + k] : 0;
}
static void gather_u64(const uint64_t* src, int A, int n, int k, uint64_t* dst) {
  for (int a = 0; a < A; ++a) dst[a] = src ? src[(size_t)a * n + k] : 0;
}
static void scatter_f1(float* dst, int A, int n, int k, const float* src) {
  for (int a = 0; a < A; ++a) dst[(size_t)a * n + k] = src[a];
}

typedef struct {
  const bmcts_config* cfg;
  const bmcts_scenario* scn;
  const bmcts_leaf_batch* in;
  const bmcts_rollout_out* out;
  int* next; /* shared work counter: threads take chunks of rollouts until the batch is drained */
} rollout_job_t;

#define ROLLOUT_CHUNK 64

static void* rollout_worker(void* arg) {
  rollout_job_t* j = (rollout_job_t*)arg;
  int A = j->scn->n_agents, n = j->in->n_leaves;
  /* unique-state form: pose / alive / depth are tables of n_states rows */
  int ns = j->in->n_states > 0 ? j->in->n_states : n;
  for (;;) {
    int begin = __atomic_fetch_add(j->next, ROLLOUT_CHUNK, __ATOMIC_RELAXED);
    if (begin >= n) break;
    int end = begin + ROLLOUT_CHUNK < n ? begin + ROLLOUT_CHUNK : n;
    for (int k = begin; k < end; ++k) {
      int row = j->in->n_states > 0 ? (int)j->in->state_index[k] : k;
      unsigned depth = j->in->depth ? j->in->depth[row] : 1u;
      float pose[MAXA * 4], ra[MAXA], fp[MAXA * 4];
      uint8_t hyp[MAXA];
      if (j->in->n_states > 0) /* the state table is state-major: record `row` is contiguous */
        memcpy(pose, j->in->pose + (size_t)row * A * 4, sizeof(float) * 4 * (size_t)A);
      else
        gather_f4(j->in->pose, A, ns, row, pose);
      gather_u8(j->in->hyp_id, A, n, k, hyp);
      oracle_rollout(j->cfg, j->scn, pose, j->in->alive[row], hyp, depth, j->in->seed,
                     j->in->stream, j->in->first_rollout_id + (uint64_t)k, &j->out->result[k], ra,
                     fp);
      if (j->out->ret_agents) scatter_f1(j->out->ret_agents, A, n, k, ra);
      if (j->out->final_pose) scatter_f4(j->out->final_pose, A, n, k, fp);
    }
  }
  return NULL;
}

int oracle_rollout_batch(const bmcts_config* cfg, const bmcts_scenario* scn,
                         const bmcts_leaf_batch* in, const bmcts_rollout_out* out,
                         int n_threads) {
  int st = check_args(cfg, scn);
  if (st) return st;
  if (!in || !out || !out->result || in->memory != BMCTS_MEM_HOST) return BMCTS_ERR_INVALID_ARG;
  if (in->n_states < 0 || (in->n_states > 0 && !in->state_index)) return BMCTS_ERR_INVALID_ARG;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  int max_useful = (in->n_leaves + ROLLOUT_CHUNK - 1) / ROLLOUT_CHUNK;
  if (n_threads > max_useful) n_threads = max_useful > 0 ? max_useful : 1;
  int next = 0;
  rollout_job_t job = {cfg, scn, in, out, &next};
  if (n_threads == 1) {
    rollout_worker(&job);
    return BMCTS_OK;
  }
  pthread_t th[256];
  for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, rollout_worker, &job);
  for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
  return BMCTS_OK;
}

/* one expansion step of batch element k (+ optional rollout from the child) */
static void step_element(const bmcts_config* cfg, const bmcts_scenario* scn,
                         const bmcts_step_in* in, const bmcts_step_out* out, int k,
                         const bmcts_rollout_out* ro, uint64_t first_rollout_id) {
  int A = scn->n_agents, n = in->n;
  unsigned depth = in->depth ? in->depth[k] : 1u;
  float pose[MAXA * 4], np[MAXA * 4], rw[MAXA], la[MAXA], c2[2];
  uint8_t hyp[MAXA], act[MAXA];
  uint64_t did[MAXA];
  uint32_t na, f;
  gather_f4(in->pose, A, n, k, pose);
  gather_u8(in->hyp_id, A, n, k, hyp);
  gather_u8(in->action, A, n, k, act);
  gather_u64(in->draw_id, A, n, k, did);
  if (act[0] == BMCTS_ACTION_NONE) {
    /* no expansion: the heuristic runs on the given (depth-capped) node itself */
    if (!ro) return;
    float ra0[MAXA], fp0[MAXA * 4];
    oracle_rollout(cfg, scn, pose, in->alive[k], hyp, depth, in->seed, in->stream,
                   first_rollout_id + (uint64_t)k, &ro->result[k], ra0, fp0);
    if (ro->ret_agents) scatter_f1(ro->ret_agents, A, n, k, ra0);
    if (ro->final_pose) scatter_f4(ro->final_pose, A, n, k, fp0);
    return;
  }
  oracle_step(cfg, scn, pose, in->alive[k], hyp, depth, act, did, in->seed, in->stream, np, &na,
              rw, c2, &f, la);
  if (out) {
    if (out->next_pose) scatter_f4(out->next_pose, A, n, k, np);
    if (out->next_alive) out->next_alive[k] = na;
    if (out->reward) scatter_f1(out->reward, A, n, k, rw);
    if (out->cost) {
      out->cost[k] = c2[0];
      out->cost[(size_t)n + k] = c2[1];
    }
    if (out->flags) out->flags[k] = f;
    if (out->last_action) scatter_f1(out->last_action, A, n, k, la);
  }
  if (!ro) return;
  bmcts_rollout_result* res = &ro->result[k];
  float ra[MAXA], fp[MAXA * 4];
  if (f & BMCTS_F_TERMINAL) {
    /* RandomHeuristic on a terminal state: zero estimate */
    memset(res, 0, sizeof(*res));
    res->flags = (f & 0xffu) | ((f & 0xffu) << 8);
    res->final_alive = na;
    memset(ra, 0, sizeof(ra));
    memcpy(fp, np, sizeof(float) * 4 * (size_t)A);
  } else {
    oracle_rollout(cfg, scn, np, na, hyp, depth + 1, in->seed, in->stream,
                   first_rollout_id + (uint64_t)k, res, ra, fp);
  }
  if (ro->ret_agents) scatter_f1(ro->ret_agents, A, n, k, ra);
  if (ro->final_pose) scatter_f4(ro->final_pose, A, n, k, fp);
}

int oracle_step_batch(const bmcts_config* cfg, const bmcts_scenario* scn,
                      const bmcts_step_in* in, const bmcts_step_out* out) {
  int st = check_args(cfg, scn);
  if (st) return st;
  if (!in || !out || in->memory != BMCTS_MEM_HOST) return BMCTS_ERR_INVALID_ARG;
  for (int k = 0; k < in->n; ++k) step_element(cfg, scn, in, out, k, NULL, 0);
  return BMCTS_OK;
}

int oracle_expand_rollout_batch(const bmcts_config* cfg, const bmcts_scenario* scn,
                                const bmcts_step_in* in, const bmcts_step_out* step_out,
                                uint64_t first_rollout_id, const bmcts_rollout_out* rollout_out,
                                int n_threads) {
  (void)n_threads;
  int st = check_args(cfg, scn);
  if (st) return st;
  if (!in || !rollout_out || !rollout_out->result || in->memory != BMCTS_MEM_HOST)
    return BMCTS_ERR_INVALID_ARG;
  for (int k = 0; k < in->n; ++k)
    step_element(cfg, scn, in, step_out, k, rollout_out, first_rollout_id);
  return BMCTS_OK;
}

/* ------------------------------------------------------------ likelihood */

int oracle_likelihood_batch(const bmcts_config* cfg, const bmcts_scenario* scn,
                            const bmcts_likelihood_in* in, float* prob) {
  /* hypothesis/idm/hypothesis_idm.cpp:40-126 */
  int st = check_args(cfg, scn);
  if (st) return st;
  if (!in || !prob || in->memory != BMCTS_MEM_HOST || in->n_buckets < 1 || in->n_samples < 1)
    return BMCTS_ERR_INVALID_ARG;
  int A = scn->n_agents, H = scn->n_hypotheses;
  float bucket_size = bm_div(in->buckets_upper - in->buckets_lower, (float)in->n_buckets); /* :76-77 */
  unsigned* hist = (unsigned*)malloc(sizeof(unsigned) * (size_t)in->n_buckets);
  if (!hist) return BMCTS_ERR_INVALID_ARG;
  for (int wi = 0; wi < in->n_worlds; ++wi) {
    world_t w;
    float lpose[MAXA * 4];
    gather_f4(in->pose, A, in->n_worlds, wi, lpose);
    load_world(scn, lpose, in->alive[wi], &w);
    for (int a = 0; a < A; ++a) {
      for (int h = 0; h < H; ++h) {
        float* out = &prob[((size_t)a * H + h) * in->n_worlds + wi];
        *out = 0.0f;
        if (!w.a[a].alive) continue; /* :53-57 agent does not exist -> 0 */
        const bmcts_hypothesis* hyp = &scn->hypotheses[h];
        /* :70-73 CalcRelativeValues in the agent's current lane corridor */
        int c = w.a[a].cur;
        float ds_lead;
        int lead = find_leader(scn, &w, a, c, &ds_lead);
        int has_leader = lead >= 0;
        float gap = 0.0f, v_l = 0.0f;
        if (has_leader) {
          gap = ds_lead - ((scn->agent_length[a] - scn->agent_rear[a]) + scn->agent_rear[lead]);
          v_l = w.a[lead].v;
        }
        memset(hist, 0, sizeof(unsigned) * (size_t)in->n_buckets);
        for (int k = 0; k < in->n_samples; ++k) { /* :82-95 */
          uint64_t id = (uint64_t)(uint32_t)k | ((uint64_t)(uint32_t)wi << 32);
          uint32_t r0[4], r1[4];
          draw(in->seed, in->stream, id, (uint32_t)a, BMCTS_RNG_LIKELIHOOD, 2u * (uint32_t)h, r0);
          draw(in->seed, in->stream, id, (uint32_t)a, BMCTS_RNG_LIKELIHOOD, 2u * (uint32_t)h + 1u,
               r1);
          idm_params_t p;
          sample_idm_params(hyp, r0, r1, &p);
          float inv_v0, inv_2sab;
          idm_constants(&p, &inv_v0, &inv_2sab);
          float acc = idm_acc(hyp, &p, inv_v0, inv_2sab, has_leader, w.a[a].v, v_l, gap);
          /* :87-91 is LOG(FATAL) in the reference; samples outside the buckets are dropped here */
          if (acc < in->buckets_lower || acc > in->buckets_upper) continue;
          int b = (int)floorf(bm_div(acc - in->buckets_lower, bucket_size));
          if (b > in->n_buckets - 1) b = in->n_buckets - 1;
          if (b < 0) b = 0;
          hist[b] += 1;
        }
        float act = in->action[(size_t)a * in->n_worlds + wi];
        if (act > in->buckets_upper || act < in->buckets_lower) continue; /* :112-116 */
        int b = (int)floorf(bm_div(act - in->buckets_lower, bucket_size));
        if (b > in->n_buckets - 1) b = in->n_buckets - 1;
        if (b < 0) b = 0;
        *out = bm_div((float)hist[b], (float)in->n_samples); /* :121-125 */
      }
    }
  }
  free(hist);
  return BMCTS_OK;
}
