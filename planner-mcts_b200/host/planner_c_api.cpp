// planner_c_api.cpp -- C ABI over the BehaviorUCT* planners (include/bmcts_planner_c.h).
#include <cmath>
#include <cstring>
#include <limits>
#include <stdexcept>

#include "planner_handle.hpp"

using namespace bark_mcts_b200;

namespace {

int Fail(bmcts_planner* p, int st, const std::string& msg) {
  if (p) p->err = msg;
  return st;
}

// ScenarioRiskFunction given by its value on every hypothesis region
class TableScenarioRiskFunction : public ScenarioRiskFunction {
 public:
  TableScenarioRiskFunction(const std::vector<BehaviorHypothesisIDM>& h, std::vector<double> v)
      : hyps_(h), values_(std::move(v)) {}
  double CalculateIntegralValue(const BehaviorHypothesisIDM& hypothesis) const override {
    for (size_t i = 0; i < hyps_.size() && i < values_.size(); ++i)
      if (std::memcmp(&hyps_[i].region, &hypothesis.region, sizeof(bmcts_hypothesis)) == 0)
        return values_[i];
    return 0.0;
  }

 private:
  std::vector<BehaviorHypothesisIDM> hyps_;
  std::vector<double> values_;
};

void EnsurePlanner(bmcts_planner* p) {
  if (p->planner) return;
  std::shared_ptr<LeafEvaluator> ev =
      p->evaluator_factory ? p->evaluator_factory()
                           : std::shared_ptr<LeafEvaluator>(MakeDefaultEvaluator(p->device));
  switch (p->kind) {
    case BMCTS_PLANNER_HYPOTHESIS:
      p->planner = std::make_shared<BehaviorUCTHypothesis>(p->params, p->hypotheses, ev);
      break;
    case BMCTS_PLANNER_RISK_CONSTRAINT:
      p->planner = std::make_shared<BehaviorUCTRiskConstraint>(
          p->params, p->hypotheses,
          p->risk_integrals.empty()
              ? nullptr
              : std::make_shared<TableScenarioRiskFunction>(p->hypotheses, p->risk_integrals),
          ev);
      break;
    case BMCTS_PLANNER_COOPERATIVE:
      p->planner = std::make_shared<BehaviorUCTCooperative>(p->params, ev);
      break;
    case BMCTS_PLANNER_NHEURISTIC_RISK_CONSTRAINT:
      if (!p->node_prior)
        throw std::runtime_error("bmcts_planner_set_node_prior was not called");
      p->planner = std::make_shared<BehaviorUCTNHeuristicRiskConstraint>(
          p->params, p->hypotheses,
          p->risk_integrals.empty()
              ? nullptr
              : std::make_shared<TableScenarioRiskFunction>(p->hypotheses, p->risk_integrals),
          p->node_prior, ev);
      break;
    default:
      throw std::runtime_error("unknown planner kind");
  }
  p->planner->set_tree_index(p->tree_index);
  if (p->evaluator_factory) {
    p->planner->set_evaluator_factory(p->evaluator_factory);
  } else {
    // one engine per tree in flight, spread over the GPUs of Mcts::Engine::Devices
    const std::vector<int> devices = p->planner->EngineDevices(p->device);
    p->planner->set_indexed_evaluator_factory([devices](int i) {
      return std::shared_ptr<LeafEvaluator>(MakeDefaultEvaluator(devices[(size_t)i % devices.size()]));
    });
  }
}

void FillResult(bmcts_planner* p, bmcts_plan_result* out) {
  std::memset(out, 0, sizeof(*out));
  BehaviorUCTBase& b = *p->planner;
  const double nan = std::numeric_limits<double>::quiet_NaN();
  for (int a = 0; a < BMCTS_MAX_EGO_ACTIONS; ++a) {
    out->return_values[a] = nan;
    out->cost_values[0][a] = nan;
    out->cost_values[1][a] = nan;
  }
  out->best_action = b.GetLastMacroAction();
  out->num_iterations = b.GetLastNumIterations();
  out->num_nodes = b.GetLastNumNodes();
  out->solution_time_ms = b.GetLastSolutionTime();
  RootStats rs = b.GetLastRootStats();
  out->num_actions = (uint32_t)rs.count.size();
  for (const auto& kv : b.GetLastReturnValues())
    if (kv.first < BMCTS_MAX_EGO_ACTIONS) out->return_values[kv.first] = kv.second;
  for (size_t a = 0; a < rs.count.size() && a < BMCTS_MAX_EGO_ACTIONS; ++a)
    out->visit_counts[a] = rs.count[a];
  if (auto* r = dynamic_cast<BehaviorUCTRiskConstraint*>(&b)) {
    auto cv = r->GetLastCostValues();
    const char* names0[] = {"envelope", "cost"};
    for (const char* nm : names0)
      if (cv.count(nm))
        for (const auto& kv : cv[nm])
          if (kv.first < BMCTS_MAX_EGO_ACTIONS) out->cost_values[0][kv.first] = kv.second;
    if (cv.count("collision"))
      for (const auto& kv : cv["collision"])
        if (kv.first < BMCTS_MAX_EGO_ACTIONS) out->cost_values[1][kv.first] = kv.second;
    for (const auto& kv : r->GetLastPolicySampled().second)
      if (kv.first < BMCTS_MAX_EGO_ACTIONS) out->policy[kv.first] = kv.second;
    auto er = r->GetLastExpectedRisk();
    auto sr = r->GetLastScenarioRisk();
    for (size_t i = 0; i < er.size() && i < 2; ++i) out->expected_risk[i] = er[i];
    for (size_t i = 0; i < sr.size() && i < 2; ++i) out->scenario_risk[i] = sr[i];
  }
  out->action_acc = (float)b.GetLastAction().acc;
  out->action_steering = (float)b.GetLastAction().steering;
  out->max_tree_depth = b.GetLastMaxTreeDepth();
  auto exp = b.GetLastRootExpandedActions();
  for (size_t i = 0; i < exp.size() && i < BMCTS_MAX_AGENTS; ++i) out->root_expanded_actions[i] = exp[i];
  Trajectory t = b.GetLastTrajectory();
  out->n_traj = (int32_t)std::min<size_t>(t.size(), BMCTS_PLAN_MAX_TRAJ);
  for (int i = 0; i < out->n_traj; ++i) {
    out->trajectory[i * 5 + 0] = t[i].t;
    out->trajectory[i * 5 + 1] = t[i].x;
    out->trajectory[i * 5 + 2] = t[i].y;
    out->trajectory[i * 5 + 3] = t[i].theta;
    out->trajectory[i * 5 + 4] = t[i].v;
  }
  out->engine_launches = b.EngineLaunchCount();
}

ObservedWorld WorldOf(const bmcts_planner* p, double world_time, uint32_t ego_id,
                             const bmcts_agent_state* agents, int n_agents) {
  ObservedWorld w;
  w.world_time = world_time;
  w.ego_id = ego_id;
  w.map = p->map;
  for (int i = 0; i < n_agents; ++i) {
    Agent a;
    a.id = agents[i].id;
    a.state = {world_time, agents[i].x, agents[i].y, agents[i].theta, agents[i].v};
    a.length = agents[i].length;
    a.width = agents[i].width;
    a.rear = agents[i].rear;
    a.last_action.acc = agents[i].last_acc;
    auto tb = p->true_behaviors.find(agents[i].id);
    if (tb != p->true_behaviors.end()) {
      a.has_behavior_model = true;
      a.behavior_model = tb->second;
    }
    auto ag = p->agent_goals.find(agents[i].id);
    if (ag != p->agent_goals.end()) {
      a.has_goal = true;
      a.goal = ag->second;
    }
    w.agents.push_back(a);
  }
  return w;
}

}  // namespace

extern "C" {

int bmcts_planner_create(int kind, int device, bmcts_planner** out) {
  if (!out || kind < BMCTS_PLANNER_HYPOTHESIS || kind > BMCTS_PLANNER_NHEURISTIC_RISK_CONSTRAINT)
    return BMCTS_ERR_INVALID_ARG;
  bmcts_planner* p = new bmcts_planner();
  p->kind = kind;
  p->device = device;
  *out = p;
  return BMCTS_OK;
}

void bmcts_planner_destroy(bmcts_planner* p) { delete p; }

int bmcts_planner_set_param(bmcts_planner* p, const char* key, const double* values, int n) {
  if (!p || !key || !values || n < 1) return BMCTS_ERR_INVALID_ARG;
  if (p->planner) return Fail(p, BMCTS_ERR_INVALID_ARG, "parameters are read at construction");
  p->params->SetListFloat(key, std::vector<double>(values, values + n));
  return BMCTS_OK;
}

int bmcts_planner_add_hypothesis(bmcts_planner* p, const bmcts_hypothesis* region, int num_samples,
                                 int num_buckets, float lower, float upper) {
  if (!p || !region || num_samples < 1 || num_buckets < 1) return BMCTS_ERR_INVALID_ARG;
  if (p->planner) return Fail(p, BMCTS_ERR_INVALID_ARG, "hypotheses are fixed at construction");
  BehaviorHypothesisIDM h;
  h.region = *region;
  h.num_samples = num_samples;
  h.num_buckets = num_buckets;
  h.buckets_lower_bound = lower;
  h.buckets_upper_bound = upper;
  p->hypotheses.push_back(h);
  return BMCTS_OK;
}

int bmcts_planner_set_scenario_risk(bmcts_planner* p, const double* v, int n) {
  if (!p || !v || n < 0) return BMCTS_ERR_INVALID_ARG;
  if (p->planner) return Fail(p, BMCTS_ERR_INVALID_ARG, "the risk function is fixed at construction");
  p->risk_integrals.assign(v, v + n);
  return BMCTS_OK;
}

int bmcts_planner_set_map(bmcts_planner* p, const bmcts_scenario* map) {
  if (!p || !map) return BMCTS_ERR_INVALID_ARG;
  p->map = std::make_shared<MapGeometry>(*map);
  return BMCTS_OK;
}

int bmcts_planner_set_tree_index(bmcts_planner* p, uint32_t tree_index) {
  if (!p) return BMCTS_ERR_INVALID_ARG;
  p->tree_index = tree_index;
  if (p->planner) p->planner->set_tree_index(tree_index);
  return BMCTS_OK;
}

int bmcts_planner_update_beliefs(bmcts_planner* p, double world_time, uint32_t ego_id,
                                 const bmcts_agent_state* agents, int n_agents) {
  if (!p || !agents || n_agents < 1) return BMCTS_ERR_INVALID_ARG;
  if (!p->map) return Fail(p, BMCTS_ERR_NO_SCENARIO, "bmcts_planner_set_map was not called");
  try {
    EnsurePlanner(p);
    auto* h = dynamic_cast<BehaviorUCTHypothesisBase*>(p->planner.get());
    if (!h) return Fail(p, BMCTS_ERR_INVALID_ARG, "this planner has no behaviour hypotheses");
    h->UpdateBeliefs(WorldOf(p, world_time, ego_id, agents, n_agents));
  } catch (const std::exception& e) {
    return Fail(p, BMCTS_ERR_CUDA, e.what());
  }
  return BMCTS_OK;
}

int bmcts_planner_reset_beliefs(bmcts_planner* p) {
  if (!p) return BMCTS_ERR_INVALID_ARG;
  if (!p->planner) return BMCTS_OK;  // nothing observed yet
  auto* h = dynamic_cast<BehaviorUCTHypothesisBase*>(p->planner.get());
  if (!h) return Fail(p, BMCTS_ERR_INVALID_ARG, "this planner has no behaviour hypotheses");
  h->ResetBeliefs();
  return BMCTS_OK;
}

int bmcts_planner_plan(bmcts_planner* p, double delta_time, double world_time, uint32_t ego_id,
                       const bmcts_agent_state* agents, int n_agents, bmcts_plan_result* out) {
  if (!p || !agents || n_agents < 1 || !out) return BMCTS_ERR_INVALID_ARG;
  if (!p->map) return Fail(p, BMCTS_ERR_NO_SCENARIO, "bmcts_planner_set_map was not called");
  try {
    EnsurePlanner(p);
    ObservedWorld w = WorldOf(p, world_time, ego_id, agents, n_agents);
    p->planner->Plan(delta_time, w);
    FillResult(p, out);
  } catch (const std::exception& e) {
    return Fail(p, BMCTS_ERR_CUDA, e.what());
  }
  return BMCTS_OK;
}

int bmcts_planner_last_edges(bmcts_planner* p, bmcts_edge_info* edges, int max_n) {
  if (!p || !p->planner) return 0;
  const auto infos = p->planner->GetLastMctsEdgeInfo();
  for (int i = 0; edges && i < max_n && i < (int)infos.size(); ++i)
    edges[i] = {infos[i].agent_id, infos[i].depth, infos[i].action, (float)infos[i].action_weight};
  return (int)infos.size();
}

int bmcts_planner_last_states(bmcts_planner* p, uint32_t* depth, uint32_t* agent_ids, float* pose,
                              int max_n, int* n_agents) {
  if (!p || !p->planner) return 0;
  const auto infos = p->planner->GetLastMctsStateInfo();
  int A = 0;
  for (const auto& s : infos) A = std::max(A, (int)s.agent_ids.size());
  if (n_agents) *n_agents = A;
  for (int i = 0; depth && agent_ids && pose && i < max_n && i < (int)infos.size(); ++i) {
    depth[i] = infos[i].depth;
    for (int a = 0; a < A; ++a) {
      const bool have = a < (int)infos[i].agent_ids.size();
      agent_ids[(size_t)i * A + a] = have ? infos[i].agent_ids[a] : 0xFFFFFFFFu;
      const State st = have ? infos[i].states[a] : State();
      float* q = pose + ((size_t)i * A + a) * 4;
      q[0] = (float)st.x;
      q[1] = (float)st.y;
      q[2] = (float)st.theta;
      q[3] = (float)st.v;
    }
  }
  return (int)infos.size();
}

int bmcts_planner_set_agent_goal(bmcts_planner* p, uint32_t agent_id, const bmcts_goal* goal) {
  if (!p) return BMCTS_ERR_INVALID_ARG;
  if (!goal)
    p->agent_goals.erase(agent_id);
  else
    p->agent_goals[agent_id] = *goal;
  return BMCTS_OK;
}

int bmcts_planner_set_true_behavior(bmcts_planner* p, uint32_t agent_id, const bmcts_hypothesis* model) {
  if (!p || !model) return BMCTS_ERR_INVALID_ARG;
  p->true_behaviors[agent_id] = *model;
  return BMCTS_OK;
}

int bmcts_planner_get_beliefs(bmcts_planner* p, uint32_t agent_id, double* beliefs, int max_h) {
  if (!p || !beliefs || !p->planner) return 0;
  auto* h = dynamic_cast<BehaviorUCTHypothesisBase*>(p->planner.get());
  if (!h) return 0;
  Beliefs b = h->GetCurrentBeliefs();
  auto it = b.find(agent_id);
  if (it == b.end()) return 0;
  int n = (int)std::min<size_t>(it->second.size(), (size_t)max_h);
  for (int i = 0; i < n; ++i) beliefs[i] = it->second[i];
  return n;
}

int bmcts_planner_root_stats(bmcts_planner* p, double* stats, int max_n) {
  if (!p || !stats || !p->planner) return 0;
  std::vector<double> f = p->planner->GetLastRootStats().Flatten();
  if ((int)f.size() > max_n) return 0;
  std::memcpy(stats, f.data(), f.size() * sizeof(double));
  return (int)f.size();
}

int bmcts_planner_decide_from_root_stats(bmcts_planner* p, const double* stats, int n,
                                         bmcts_plan_result* out) {
  if (!p || !stats || !out || !p->planner) return BMCTS_ERR_INVALID_ARG;
  RootStats rs = p->planner->GetLastRootStats();
  if (n != (int)rs.Flatten().size()) return Fail(p, BMCTS_ERR_INVALID_ARG, "root stats size mismatch");
  rs.Unflatten(std::vector<double>(stats, stats + n));
  try {
    p->planner->DecideFromRootStats(rs);
    p->planner->RefreshTrajectory();  // of the action chosen from the merged statistics
    FillResult(p, out);
  } catch (const std::exception& e) {
    return Fail(p, BMCTS_ERR_CUDA, e.what());
  }
  return BMCTS_OK;
}

int bmcts_planner_set_node_prior(bmcts_planner* p, bmcts_node_prior_fn fn, void* user) {
  if (!p || !fn) return BMCTS_ERR_INVALID_ARG;
  if (p->kind != BMCTS_PLANNER_NHEURISTIC_RISK_CONSTRAINT)
    return Fail(p, BMCTS_ERR_INVALID_ARG, "only the NHeuristic risk-constraint planner takes a node prior");
  if (p->planner) return Fail(p, BMCTS_ERR_INVALID_ARG, "the node prior is fixed at construction");
  p->node_prior = std::make_shared<CallbackPriorProvider>(fn, user);
  return BMCTS_OK;
}

int bmcts_planner_observation_size(bmcts_planner* p) {
  if (!p) return BMCTS_ERR_INVALID_ARG;
  return NearestObserver::FromParams(*p->params).observation_size();
}

int bmcts_planner_observe(bmcts_planner* p, uint32_t ego_id, const bmcts_agent_state* agents,
                          int n_agents, float* observation, int max_n) {
  if (!p || !agents || n_agents < 1 || n_agents > BMCTS_MAX_AGENTS || !observation)
    return BMCTS_ERR_INVALID_ARG;
  if (!p->map) return Fail(p, BMCTS_ERR_NO_SCENARIO, "bmcts_planner_set_map was not called");
  NearestObserver o = NearestObserver::FromParams(*p->params);
  o.SetWorldBounds(*p->map);
  if (o.observation_size() > max_n) return Fail(p, BMCTS_ERR_INVALID_ARG, "observation buffer too small");
  // slot 0 = ego, the others in the order given
  std::vector<float> pose((size_t)n_agents * 4, 0.0f);
  int slot = 1, found = 0;
  for (int i = 0; i < n_agents; ++i) {
    const int s = agents[i].id == ego_id && !found ? 0 : slot++;
    if (s == 0) found = 1;
    if (s >= n_agents) return Fail(p, BMCTS_ERR_INVALID_ARG, "ego agent not in the world");
    pose[s * 4 + 0] = agents[i].x;
    pose[s * 4 + 1] = agents[i].y;
    pose[s * 4 + 2] = agents[i].theta;
    pose[s * 4 + 3] = agents[i].v;
  }
  if (!found) return Fail(p, BMCTS_ERR_INVALID_ARG, "ego agent not in the world");
  o.Observe(pose.data(), n_agents >= 32 ? 0xffffffffu : ((1u << n_agents) - 1u), n_agents, observation);
  return o.observation_size();
}

int bmcts_planner_node_prior_stats(bmcts_planner* p, uint64_t* calls, uint64_t* nodes) {
  if (!p || !p->node_prior) return BMCTS_ERR_INVALID_ARG;
  if (calls) *calls = p->node_prior->calls();
  if (nodes) *nodes = p->node_prior->nodes();
  return BMCTS_OK;
}

double bmcts_planner_parameter_sampling_probability(bmcts_planner* p, int h) {
  if (!p || h < 0 || h >= (int)p->hypotheses.size()) return 0.0;
  return p->hypotheses[h].ParameterSamplingProbability();
}

const char* bmcts_planner_last_error(const bmcts_planner* p) { return p ? p->err.c_str() : "null planner"; }

}  // extern "C"
