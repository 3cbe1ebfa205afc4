// scenario.cpp -- host-side parts of the C ABI that need no GPU:
// reference parameter defaults and scenario validation / derived geometry.
//
// Paths cited are relative to /root/reference/bark_mcts/models/behavior/.
#include <cstring>

#include "bmcts_c.h"
#include "bmcts_math.h"

extern "C" {

int bmcts_abi_version(void) { return BMCTS_ABI_VERSION; }

// Defaults of param_config/mcts_state_parameters_from_param_server.hpp:19-51,
// param_config/mcts_parameters_from_param_server.hpp:45,55-58 and the
// evaluator / ego-model subtree of tests/data/nheuristic_test.json:152-199,281-311.
void bmcts_config_default(bmcts_config* c) {
  if (!c) return;
  std::memset(c, 0, sizeof(*c));
  c->goal_reward = 100.0f;
  c->collision_reward = -1000.0f;
  c->safe_dist_violated_reward = -0.1f;
  c->out_of_drivable_reward = 0.0f;
  c->goal_cost = -100.0f;
  c->collision_cost = 1000.0f;
  c->safe_dist_violated_cost = 0.1f;
  c->out_of_drivable_cost = 0.0f;
  c->cooperation_factor = 0.2f;
  c->step_reward = 0.0f;
  c->prediction_k = 0.5f;
  c->prediction_alpha = 0.0f;
  c->normalization_tau = 0.2f;
  c->split_safe_dist_collision = 0;
  c->chance_costs = 0;
  c->add_safe_dist = 0;
  c->static_safe_dist_is_terminal = 0;
  c->dynamic_safe_dist_is_terminal = 0;
  c->out_of_drivable_is_terminal = 0;
  c->drivable_lat_safe_dist = 0.5f;
  c->drivable_lon_safe_dist = 0.5f;
  c->static_lat_safe_dist = 1.5f;
  c->static_lon_safe_dist = 1.5f;
  c->dyn_max_other_decel = 5.0f;
  c->dyn_max_ego_decel = 5.0f;
  c->dyn_reaction_time_others = 0.8f;
  c->dyn_reaction_time_ego = 0.2f;
  c->dyn_to_rear = 1;
  c->discount_factor = 0.9f;
  c->max_rollout_steps = 1000;
  c->wheel_base = 2.7f;
  c->delta_max = 0.2f;
  c->crosstrack_gain = 1.0f;
  c->lon_acc_min = -8.0f;
  c->lon_acc_max = 4.0f;
  c->ego_max_velocity = 50.0f;
  c->num_substeps = 10;
  c->state_kind = BMCTS_STATE_HYPOTHESIS;
  c->limit_steering_rate = 1;      // tests/data/nheuristic_test.json:183-186
  c->lat_acc_max = 4.0f;           // :191-196
  c->lat_acc_min = -4.0f;
  c->ego_integration_dt = 0.02f;   // :153-155 (0.019999999552965164 = float(0.02))
}

int bmcts_config_validate(const bmcts_config* c) {
  if (!c) return BMCTS_ERR_INVALID_ARG;
  if (c->num_substeps < 1 || c->num_substeps > 100) return BMCTS_ERR_INVALID_ARG;
  if (c->max_rollout_steps < 0 || c->max_rollout_steps >= BMCTS_MAX_DEPTH - 64)
    return BMCTS_ERR_INVALID_ARG;
  if (!(c->wheel_base > 0.0f)) return BMCTS_ERR_INVALID_ARG;
  if (!(c->delta_max > 0.0f) || c->delta_max > BM_PIO4) return BMCTS_ERR_INVALID_ARG;
  if (!(c->prediction_k > 0.0f)) return BMCTS_ERR_INVALID_ARG;
  if (!(c->dyn_max_other_decel > 0.0f) || !(c->dyn_max_ego_decel > 0.0f))
    return BMCTS_ERR_INVALID_ARG;
  if (c->limit_steering_rate && !(c->lat_acc_min <= 0.0f && c->lat_acc_max >= 0.0f))
    return BMCTS_ERR_INVALID_ARG;
  // at most 1024 Euler steps per macro action at the shortest prediction span
  if (c->ego_integration_dt > 0.0f && c->prediction_k / c->ego_integration_dt > 1024.0f)
    return BMCTS_ERR_INVALID_ARG;
  if (c->state_kind < BMCTS_STATE_HYPOTHESIS || c->state_kind > BMCTS_STATE_COOPERATIVE)
    return BMCTS_ERR_INVALID_ARG;
  return BMCTS_OK;
}

int bmcts_scenario_finalize(bmcts_scenario* s) {
  if (!s) return BMCTS_ERR_INVALID_ARG;
  s->finalized = 0;
  if (s->n_corridors < 1 || s->n_corridors > BMCTS_MAX_CORRIDORS) return BMCTS_ERR_INVALID_ARG;
  if (s->n_road_pts < 3 || s->n_road_pts > BMCTS_MAX_ROAD_PTS) return BMCTS_ERR_INVALID_ARG;
  if (s->n_agents < 1 || s->n_agents > BMCTS_MAX_AGENTS) return BMCTS_ERR_INVALID_ARG;
  if (s->n_hypotheses < 0 || s->n_hypotheses > BMCTS_MAX_HYPOTHESES) return BMCTS_ERR_INVALID_ARG;
  if (s->n_ego_actions < 1 || s->n_ego_actions > BMCTS_MAX_EGO_ACTIONS)
    return BMCTS_ERR_INVALID_ARG;
  for (int c = 0; c < s->n_corridors; ++c) {
    bmcts_corridor* k = &s->corridors[c];
    if (k->n_pts < 2 || k->n_pts > BMCTS_MAX_CORRIDOR_PTS) return BMCTS_ERR_INVALID_ARG;
    if (!(k->half_width > 0.0f)) return BMCTS_ERR_INVALID_ARG;
    if (k->left < -1 || k->left >= s->n_corridors || k->right < -1 || k->right >= s->n_corridors)
      return BMCTS_ERR_INVALID_ARG;
    float acc = 0.0f;
    for (int i = 0; i < BMCTS_MAX_CORRIDOR_PTS; ++i) {
      k->tx[i] = k->ty[i] = k->seg_len[i] = k->s0[i] = k->angle[i] = 0.0f;
    }
    for (int i = 0; i + 1 < k->n_pts; ++i) {
      float dx = k->px[i + 1] - k->px[i];
      float dy = k->py[i + 1] - k->py[i];
      float len = bm_sqrt(bm_fma(dx, dx, dy * dy));
      if (!(len > 1.0e-4f)) return BMCTS_ERR_INVALID_ARG;
      k->tx[i] = bm_div(dx, len);
      k->ty[i] = bm_div(dy, len);
      k->seg_len[i] = len;
      k->s0[i] = acc;
      k->angle[i] = bm_atan2(dy, dx);
      acc = acc + len;
    }
    k->length = acc;
  }
  for (int a = 0; a < s->n_agents; ++a) {
    if (!(s->agent_length[a] > 0.0f) || !(s->agent_width[a] > 0.0f)) return BMCTS_ERR_INVALID_ARG;
    if (s->agent_rear[a] < 0.0f || s->agent_rear[a] > s->agent_length[a])
      return BMCTS_ERR_INVALID_ARG;
  }
  for (int h = 0; h < s->n_hypotheses; ++h) {
    const bmcts_hypothesis* hy = &s->hypotheses[h];
    for (int m = 0; m < BMCTS_NUM_IDM_PARAMS; ++m)
      if (!(hy->lo[m] <= hy->hi[m])) return BMCTS_ERR_INVALID_ARG;
    // the CAH / ACC blend of BARK's coolness factor is out of scope;
    // every hypothesis in the reference's tests and parameter files fixes it at 0
    if (hy->lo[BMCTS_IDM_COOLNESS] != 0.0f || hy->hi[BMCTS_IDM_COOLNESS] != 0.0f)
      return BMCTS_ERR_UNSUPPORTED;
    if (hy->exponent < 0 || hy->exponent > 16) return BMCTS_ERR_INVALID_ARG;
    if (!(hy->acc_lower_bound <= hy->acc_upper_bound)) return BMCTS_ERR_INVALID_ARG;
    if (!(hy->min_velocity <= hy->max_velocity)) return BMCTS_ERR_INVALID_ARG;
  }
  for (int a = 0; a < s->n_ego_actions; ++a) {
    int t = s->ego_actions[a].type;
    if (t < BMCTS_ACT_CONST_ACC_STAY_LANE || t > BMCTS_ACT_CHANGE_RIGHT)
      return BMCTS_ERR_INVALID_ARG;
  }
  auto goal_ok = [&](const bmcts_goal& g) {
    if (g.kind == BMCTS_GOAL_POLYGON) return g.n_pts >= 3 && g.n_pts <= BMCTS_MAX_GOAL_PTS;
    if (g.kind == BMCTS_GOAL_FRENET_LIMITS) return g.corridor >= 0 && g.corridor < s->n_corridors;
    return g.kind == BMCTS_GOAL_NONE;
  };
  if (!goal_ok(s->goal)) return BMCTS_ERR_INVALID_ARG;
  if (s->agent_goal[0] != 0) return BMCTS_ERR_INVALID_ARG;  // the ego's goal is `goal`
  for (int a = 0; a < s->n_agents; ++a) {
    const int k = s->agent_goal[a];
    if (k < 0 || k > BMCTS_MAX_EXTRA_GOALS) return BMCTS_ERR_INVALID_ARG;
    if (k > 0 && !goal_ok(s->extra_goals[k - 1])) return BMCTS_ERR_INVALID_ARG;
  }
  s->finalized = 1;
  return BMCTS_OK;
}

// sizes of the ABI structs, so that bindings (ctypes, cgo, JNI ...) can verify their mirror
size_t bmcts_struct_size(int which) {
  switch (which) {
    case 0: return sizeof(bmcts_config);
    case 1: return sizeof(bmcts_corridor);
    case 2: return sizeof(bmcts_hypothesis);
    case 3: return sizeof(bmcts_ego_action);
    case 4: return sizeof(bmcts_goal);
    case 5: return sizeof(bmcts_scenario);
    case 6: return sizeof(bmcts_rollout_result);
    case 7: return sizeof(bmcts_leaf_batch);
    case 8: return sizeof(bmcts_rollout_out);
    case 9: return sizeof(bmcts_step_in);
    case 10: return sizeof(bmcts_step_out);
    case 11: return sizeof(bmcts_likelihood_in);
    default: return 0;
  }
}

const char* bmcts_status_string(int st) {
  switch (st) {
    case BMCTS_OK: return "ok";
    case BMCTS_ERR_INVALID_ARG: return "invalid argument";
    case BMCTS_ERR_CUDA: return "CUDA error";
    case BMCTS_ERR_NO_SCENARIO: return "no scenario set";
    case BMCTS_ERR_UNSUPPORTED: return "unsupported feature";
    case BMCTS_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
    case BMCTS_ERR_COMM: return "communicator error";
    default: return "unknown status";
  }
}

}  // extern "C"
