// behavior_uct.cpp -- BehaviorUCT* planners (see behavior_uct.hpp).
// Paths cited are relative to /root/reference/bark_mcts/models/behavior/.
#include "behavior_uct.hpp"

#include <atomic>
#include <thread>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <stdexcept>

#include "bmcts_math.h"

namespace bark_mcts_b200 {

namespace {

const char* kBase = "BehaviorUctBase::";

struct HostProj {
  float s, d;
  int seg;
  bool in;
};

// host restatement of the engine's Frenet projection; used only outside the hot path
// (valid macro actions at the root, trajectory of the chosen action)
HostProj ProjectOn(const bmcts_corridor& c, float x, float y) {
  const int nseg = c.n_pts - 1;
  float best_d2 = 1.0e30f, best_u = 0, best_raw = 0, best_cross = 0;
  int best = 0;
  for (int k = 0; k < nseg; ++k) {
    float rx = x - c.px[k], ry = y - c.py[k];
    float u = bm_fma(rx, c.tx[k], ry * c.ty[k]);
    float uc = bm_clamp(u, 0.0f, c.seg_len[k]);
    float ex = bm_fma(-uc, c.tx[k], rx), ey = bm_fma(-uc, c.ty[k], ry);
    float d2 = bm_fma(ex, ex, ey * ey);
    if (d2 < best_d2) {
      best_d2 = d2;
      best = k;
      best_u = uc;
      best_raw = u;
      best_cross = bm_fma(c.tx[k], ry, -(c.ty[k] * rx));
    }
  }
  HostProj p{c.s0[best] + best_u, best_cross, best, true};
  if (best == 0 && best_raw < 0.0f) p.in = false;
  if (best == nseg - 1 && best_raw > c.seg_len[best]) p.in = false;
  return p;
}

int CurrentCorridor(const bmcts_scenario& scn, float x, float y) {
  float best = 1.0e30f;
  int cur = 0;
  for (int c = 0; c < scn.n_corridors; ++c) {
    HostProj p = ProjectOn(scn.corridors[c], x, y);
    float m = p.in ? std::fabs(p.d) : 1.0e30f;
    if (m < best) {
      best = m;
      cur = c;
    }
  }
  return cur;
}

}  // namespace

// ------------------------------------------------------------------ BehaviorUCTBase

BehaviorUCTBase::BehaviorUCTBase(const ParamsPtr& params, int state_kind, int device)
    : BehaviorUCTBase(params, state_kind, std::shared_ptr<LeafEvaluator>(MakeDefaultEvaluator(device))) {
  const std::vector<int> devices = EngineDevices(device);
  evaluator_factory_ = [devices](int i) {
    return std::shared_ptr<LeafEvaluator>(MakeDefaultEvaluator(devices[(size_t)i % devices.size()]));
  };
}

BehaviorUCTBase::BehaviorUCTBase(const ParamsPtr& params, int state_kind,
                                 std::shared_ptr<LeafEvaluator> evaluator)
    : params_(params ? params : std::make_shared<Params>()),
      state_kind_(state_kind),
      evaluator_(std::move(evaluator)),
      mcts_parameters_(MctsParametersFromParamServer(*params_, kBase)),
      engine_config_(EngineConfigFromParamServer(*params_, kBase, state_kind, mcts_parameters_)),
      max_nearest_agents_((unsigned)params_->GetInt(std::string(kBase) + "MaxNearestAgents", 5)),
      constant_action_idx_((int)params_->GetInt(std::string(kBase) + "ConstantActionIndex", -1)),
      extract_edge_info_(params_->GetBool(std::string(kBase) + "ExtractEdgeInfo", true)),
      extract_state_info_(params_->GetBool(std::string(kBase) + "ExtractStateInfo", false)),
      max_extraction_depth_((unsigned)params_->GetInt(std::string(kBase) + "MaxExtractionDepth", 10)) {
  if (!evaluator_) throw std::runtime_error("BehaviorUCT*: no rollout engine (no CPU fallback)");
}

std::string BehaviorUCTBase::GetPrimitiveName(unsigned action) const {
  (void)action;
  return "MacroAction";
}

ObservedWorld BehaviorUCTBase::FilterAgents(const ObservedWorld& w) const {
  const Agent* ego = w.GetEgoAgent();
  ObservedWorld out = w;
  if (!ego) return out;
  std::vector<std::pair<double, const Agent*>> by_dist;
  for (const Agent& a : w.agents) {
    if (a.id == w.ego_id) continue;
    const double dx = a.state.x - ego->state.x, dy = a.state.y - ego->state.y;
    by_dist.push_back({dx * dx + dy * dy, &a});
  }
  std::sort(by_dist.begin(), by_dist.end(),
            [](const std::pair<double, const Agent*>& l, const std::pair<double, const Agent*>& r) {
              return l.first < r.first || (l.first == r.first && l.second->id < r.second->id);
            });
  out.agents.clear();
  out.agents.push_back(*ego);
  for (size_t i = 0; i < by_dist.size() && i < max_nearest_agents_; ++i)
    out.agents.push_back(*by_dist[i].second);
  return out;
}

BehaviorUCTBase::SearchWorld BehaviorUCTBase::BuildSearchWorld(
    const ObservedWorld& world, const std::vector<BehaviorHypothesisIDM>& hypotheses) const {
  if (!world.map) throw std::runtime_error("ObservedWorld without map geometry");
  SearchWorld sw;
  sw.scenario = *world.map;
  bmcts_scenario& s = sw.scenario;
  // slot order: ego first, others by ascending AgentId (mcts_state_base.hpp:218-226)
  std::vector<const Agent*> order;
  const Agent* ego = world.GetEgoAgent();
  if (ego) order.push_back(ego);
  std::vector<const Agent*> others;
  for (const Agent& a : world.agents)
    if (a.id != world.ego_id) others.push_back(&a);
  std::sort(others.begin(), others.end(), [](const Agent* l, const Agent* r) { return l->id < r->id; });
  for (const Agent* a : others)
    if ((int)order.size() < BMCTS_MAX_AGENTS) order.push_back(a);
  if (order.empty()) throw std::runtime_error("ObservedWorld without agents");
  s.n_agents = (int)order.size();
  sw.pose.assign((size_t)s.n_agents * 4, 0.0f);
  for (int i = 0; i < s.n_agents; ++i) {
    const Agent* a = order[i];
    sw.slot_ids.push_back(a->id);
    s.agent_length[i] = a->length;
    s.agent_width[i] = a->width;
    s.agent_rear[i] = a->rear;
    sw.pose[(size_t)i * 4 + 0] = (float)a->state.x;
    sw.pose[(size_t)i * 4 + 1] = (float)a->state.y;
    sw.pose[(size_t)i * 4 + 2] = (float)a->state.theta;
    sw.pose[(size_t)i * 4 + 3] = (float)a->state.v;
  }
  // goals of the other agents (cooperative state): distinct definitions go to extra_goals
  int n_extra = 0;
  for (int i = 0; i < BMCTS_MAX_AGENTS; ++i) s.agent_goal[i] = 0;
  for (int i = 1; i < s.n_agents; ++i) {
    const Agent* a = order[i];
    if (!a->has_goal || std::memcmp(&a->goal, &s.goal, sizeof(bmcts_goal)) == 0) continue;
    int k = 0;
    while (k < n_extra && std::memcmp(&a->goal, &s.extra_goals[k], sizeof(bmcts_goal)) != 0) ++k;
    if (k == n_extra) {
      if (n_extra == BMCTS_MAX_EXTRA_GOALS)
        throw std::runtime_error("more distinct agent goals than the engine holds (1 + " +
                                 std::to_string(BMCTS_MAX_EXTRA_GOALS) + ")");
      s.extra_goals[n_extra++] = a->goal;
    }
    s.agent_goal[i] = k + 1;
  }
  sw.alive = ego ? ((s.n_agents >= 32) ? 0xffffffffu : ((1u << s.n_agents) - 1u)) : 0u;
  s.n_hypotheses = (int)std::min<size_t>(hypotheses.size(), BMCTS_MAX_HYPOTHESES);
  for (int h = 0; h < s.n_hypotheses; ++h) s.hypotheses[h] = hypotheses[h].region;
  // BehaviorMacroActionsFromParamServer: const-acc primitives, then the lane changes whose
  // precondition holds at the root (GetNumMotionPrimitives, behavior_uct_hypothesis.cpp:40-41)
  const std::string eg = std::string(kBase) + "EgoBehavior::";
  std::vector<double> accs = params_->GetListFloat(eg + "AccelerationInputs", {0.0, 1.0, 4.0, -1.0, -8.0});
  const bool lane_changes = params_->GetBool(eg + "AddLaneChangeActions", true);
  s.n_ego_actions = 0;
  for (double a : accs)
    if (s.n_ego_actions < BMCTS_MAX_EGO_ACTIONS - 2)
      s.ego_actions[s.n_ego_actions++] = {BMCTS_ACT_CONST_ACC_STAY_LANE, (float)a};
  if (s.n_ego_actions == 0) s.ego_actions[s.n_ego_actions++] = {BMCTS_ACT_CONST_ACC_STAY_LANE, 0.0f};
  int st = bmcts_scenario_finalize(&s);
  if (st != BMCTS_OK)
    throw std::runtime_error(std::string("bmcts_scenario_finalize: ") + bmcts_status_string(st));
  if (lane_changes && ego) {
    const int cur = CurrentCorridor(s, sw.pose[0], sw.pose[1]);
    if (s.corridors[cur].left >= 0) s.ego_actions[s.n_ego_actions++] = {BMCTS_ACT_CHANGE_LEFT, 0.0f};
    if (s.corridors[cur].right >= 0) s.ego_actions[s.n_ego_actions++] = {BMCTS_ACT_CHANGE_RIGHT, 0.0f};
  }
  return sw;
}

Trajectory BehaviorUCTBase::MacroActionTrajectory(const SearchWorld& sw, unsigned action,
                                                  double delta_time, double t0,
                                                  Action* last_action) const {
  // BehaviorMPMacroActions::Plan of the chosen primitive from the current ego state; the same
  // single-track / Stanley update as the engine's macro branch
  const bmcts_scenario& scn = sw.scenario;
  const bmcts_config& cfg = engine_config_;
  const bmcts_ego_action act = scn.ego_actions[std::min<unsigned>(action, scn.n_ego_actions - 1)];
  const float acc = bm_clamp(act.acceleration, cfg.lon_acc_min, cfg.lon_acc_max);
  float x = sw.pose[0], y = sw.pose[1], th = sw.pose[2], v = sw.pose[3];
  int tgt = CurrentCorridor(scn, x, y);
  if (act.type == BMCTS_ACT_CHANGE_LEFT && scn.corridors[tgt].left >= 0) tgt = scn.corridors[tgt].left;
  else if (act.type == BMCTS_ACT_CHANGE_RIGHT && scn.corridors[tgt].right >= 0) tgt = scn.corridors[tgt].right;
  const bmcts_corridor& corr = scn.corridors[tgt];
  // BehaviorMotionPrimitives::IntegrationTimeDelta: ceil(delta_time / delta) Euler steps
  int nsub = cfg.num_substeps;
  if (cfg.ego_integration_dt > 0.0f)
    nsub = (int)bm_clamp(std::ceil(bm_div((float)delta_time, cfg.ego_integration_dt) - 1.0e-3f), 1.0f, 1024.0f);
  const float dt = bm_div((float)delta_time, (float)nsub);
  const float inv_wb = bm_div(1.0f, cfg.wheel_base);
  Trajectory traj;
  traj.push_back({t0, x, y, th, v});
  float delta = 0.0f;
  for (int k = 0; k < nsub; ++k) {
    float sn, cs;
    bm_sincos(th, &sn, &cs);
    float fx = bm_fma(cfg.wheel_base, cs, x), fy = bm_fma(cfg.wheel_base, sn, y);
    HostProj pr = ProjectOn(corr, fx, fy);
    float th_err = bm_norm_angle(corr.angle[pr.seg] - th);
    delta = th_err + bm_atan2(-(cfg.crosstrack_gain * pr.d), v);
    delta = bm_clamp(delta, -cfg.delta_max, cfg.delta_max);
    float tn = bm_tan_small(delta);
    if (cfg.limit_steering_rate) {
      const float r = bm_div(cfg.wheel_base, bm_max(v * v, 1.0e-6f));
      tn = bm_clamp(tn, cfg.lat_acc_min * r, cfg.lat_acc_max * r);
    }
    float dv = dt * v;
    x = bm_fma(dv, cs, x);
    y = bm_fma(dv, sn, y);
    th = bm_fma(dv, tn * inv_wb, th);
    v = bm_clamp(bm_fma(dt, acc, v), 0.0f, cfg.ego_max_velocity);
    traj.push_back({t0 + (k + 1) * (double)dt, x, y, k + 1 == nsub ? bm_norm_angle(th) : th, v});
  }
  if (last_action) *last_action = {acc, delta};
  return traj;
}

void BehaviorUCTBase::FinishPlan(const SearchWorld& sw, const ObservedWorld& observed_world,
                                 unsigned best_action, double min_planning_time) {
  last_motion_idx_ = best_action;
  if (&sw != &last_sw_) {
    last_sw_ = sw;
    last_plan_world_.ego_id = observed_world.ego_id;
    last_plan_world_.world_time = observed_world.world_time;
    last_plan_time_ = min_planning_time;
    has_last_plan_ = true;
  }
  const unsigned applied = constant_action_idx_ >= 0 ? (unsigned)constant_action_idx_ : best_action;
  last_trajectory_ = MacroActionTrajectory(sw, applied, min_planning_time, observed_world.world_time,
                                           &last_action_);
}

void BehaviorUCTBase::DecideFromRootStats(const RootStats& merged) {
  // returnBestAction(): highest mean return among the visited root actions
  last_root_stats_ = merged;
  Policy pol;
  unsigned best = 0;
  double best_v = -1e300;
  for (size_t a = 0; a < merged.count.size(); ++a) {
    if (merged.count[a] <= 0) continue;
    const double v = merged.ret_sum[a] / merged.count[a];
    pol[(unsigned)a] = v;
    if (v > best_v) {
      best_v = v;
      best = (unsigned)a;
    }
  }
  last_return_values_ = pol;
  last_motion_idx_ = best;
  last_num_iterations_ = (unsigned)merged.iterations;
  last_num_nodes_ = (unsigned)merged.nodes;
}

BehaviorUCTBase::TreeResult BehaviorUCTBase::CollectTree(Mcts& mcts) {
  TreeResult r;
  r.stats = mcts.GetRootStats();
  r.time_ms = mcts.searchTime();
  r.max_depth = mcts.maxTreeDepth();
  r.root_expanded = mcts.RootExpandedActions();
  return r;
}

RootStats BehaviorUCTBase::RunTrees(const SearchWorld& sw, long trees, const StartTree& start_tree,
                                    const std::function<TreeResult(Mcts&)>& collect,
                                    std::vector<double>* mean_lambdas) {
  std::vector<TreeResult> results((size_t)trees);
  const auto t_begin = std::chrono::steady_clock::now();
  long n_threads = 1;
  if (mcts_parameters_.USE_MULTI_THREADING && trees > 1 && evaluator_factory_) {
    n_threads = std::min<long>(trees, std::max(1u, std::thread::hardware_concurrency()));
    if (mcts_parameters_.MAX_THREADS > 0) n_threads = std::min(n_threads, mcts_parameters_.MAX_THREADS);
    n_threads = std::min<long>(n_threads, 64);
  }
  // A thread keeps the waves of up to `lanes` trees in flight (one engine = one CUDA stream per
  // tree): while the GPU rolls out the wave of one tree the thread selects / backs up another.
  // Trees are still independent searches: what a tree computes does not depend on the schedule.
  long lanes = 1;
  if (evaluator_factory_ && trees > n_threads)
    lanes = std::min<long>(kMaxTreesInFlight, (trees + n_threads - 1) / n_threads);
  const bool own_engines = evaluator_factory_ && (n_threads > 1 || lanes > 1);
  if (own_engines) {
    // engines are kept between Plan() calls
    while ((long)tree_evaluators_.size() < n_threads * lanes)
      tree_evaluators_.push_back(evaluator_factory_((int)tree_evaluators_.size()));
    for (long i = 0; i < n_threads * lanes; ++i) {
      int st = tree_evaluators_[i]->Configure(engine_config_, sw.scenario);
      if (st != BMCTS_OK)
        throw std::runtime_error("engine configure failed: " + tree_evaluators_[i]->LastError());
    }
  }
  std::atomic<long> next(0);
  struct Slot {
    long tree = -1;
    std::unique_ptr<Mcts> mcts;
    LeafEvaluator* evaluator = nullptr;
  };
  // runs trees from the shared counter on this thread's engines until none is left
  // the tree infos are taken from the last tree only (they describe ONE tree)
  auto collect_tree = [&](long t, Mcts& mcts) {
    TreeResult r = collect(mcts);
    if (t == trees - 1) {
      if (extract_edge_info_) r.edges = mcts.ExtractEdges(max_extraction_depth_);
      if (extract_state_info_) r.states = mcts.ExtractStates(max_extraction_depth_);
    }
    return r;
  };
  auto worker = [&](long thread_index) {
    std::vector<Slot> slots((size_t)lanes);
    for (long k = 0; k < lanes; ++k)
      slots[k].evaluator = own_engines ? tree_evaluators_[thread_index * lanes + k].get() : evaluator_.get();
    // next tree into the slot, first wave on its way; false when no tree is left
    auto refill = [&](Slot& s) {
      for (;;) {
        const long t = next.fetch_add(1);
        if (t >= trees) {
          s.tree = -1;
          s.mcts.reset();
          return false;
        }
        s.tree = t;
        s.mcts = start_tree(t, s.evaluator);
        if (!s.mcts->Done()) {
          s.mcts->BeginWave();
          return true;
        }
        s.mcts->FinishSearch();  // nothing to do (zero budget)
        results[t] = collect_tree(t, *s.mcts);
      }
    };
    long active = 0;
    try {
      for (Slot& s : slots) active += refill(s);
      while (active > 0) {
        for (Slot& s : slots) {
          if (s.tree < 0) continue;
          s.mcts->EndWave();
          if (!s.mcts->Done()) {
            s.mcts->BeginWave();
            continue;
          }
          s.mcts->FinishSearch();
          results[s.tree] = collect_tree(s.tree, *s.mcts);
          if (!refill(s)) active--;
        }
      }
    } catch (...) {
      // leave no wave in flight behind: the engines are reused by the next Plan()
      for (Slot& s : slots)
        if (s.evaluator) s.evaluator->EndExpandRollout();
      throw;
    }
  };
  if (n_threads <= 1) {
    worker(0);
  } else {
    std::vector<std::string> errors((size_t)n_threads);
    std::vector<std::thread> pool;
    for (long i = 0; i < n_threads; ++i)
      pool.emplace_back([&, i]() {
        try {
          worker(i);
        } catch (const std::exception& ex) {
          errors[i] = ex.what();
        }
      });
    for (auto& th : pool) th.join();
    for (const auto& err : errors)
      if (!err.empty()) throw std::runtime_error(err);
  }
  RootStats merged;
  merged.Resize(sw.scenario.n_ego_actions);
  // trees overlap in time (threads, interleaved waves): the solution time is the wall clock
  const double time_ms =
      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
  if (mean_lambdas) mean_lambdas->clear();
  for (long t = 0; t < trees; ++t) {  // merge_node_statistics, in tree order (deterministic)
    merged.Add(results[t].stats);
    if (mean_lambdas) {  // merge_mcts_parameters: mean lambda
      if (mean_lambdas->size() < results[t].lambdas.size()) mean_lambdas->resize(results[t].lambdas.size(), 0.0);
      for (size_t i = 0; i < results[t].lambdas.size(); ++i)
        (*mean_lambdas)[i] += results[t].lambdas[i] / (double)trees;
    }
  }
  last_solution_time_ = time_ms;
  last_max_tree_depth_ = results[trees - 1].max_depth;
  last_root_expanded_actions_ = results[trees - 1].root_expanded;
  // ExtractMctsEdgeInfo / ExtractMctsStateInfo of the last tree (behavior_uct_hypothesis.cpp:69-79)
  mcts_edge_infos_.clear();
  for (const Mcts::EdgeInfo& e : results[trees - 1].edges)
    mcts_edge_infos_.push_back({sw.slot_ids[(size_t)e.slot], e.depth, e.action, e.weight});
  mcts_state_infos_.clear();
  for (const Mcts::StateInfo& st : results[trees - 1].states) {
    MctsStateInfo info;
    info.depth = st.depth;
    for (size_t a = 0; a < sw.slot_ids.size(); ++a) {
      if (!((st.alive >> a) & 1u)) continue;
      info.agent_ids.push_back(sw.slot_ids[a]);
      info.states.push_back({0.0, st.pose[a * 4], st.pose[a * 4 + 1], st.pose[a * 4 + 2], st.pose[a * 4 + 3]});
    }
    mcts_state_infos_.push_back(std::move(info));
  }
  return merged;
}

// ------------------------------------------------------- BehaviorUCTHypothesisBase

BehaviorUCTHypothesisBase::BehaviorUCTHypothesisBase(const ParamsPtr& params,
                                                     const std::vector<BehaviorHypothesisIDM>& h,
                                                     int state_kind, int device)
    : BehaviorUCTBase(params, state_kind, device),
      behavior_hypotheses_(h),
      use_true_behaviors_as_hypothesis_(params_->GetBool(
          "BehaviorUctHypothesis::PredictionSettings::UseTrueBehaviorsAsHypothesis", false)),
      belief_tracker_(mcts_parameters_) {}

BehaviorUCTHypothesisBase::BehaviorUCTHypothesisBase(const ParamsPtr& params,
                                                     const std::vector<BehaviorHypothesisIDM>& h,
                                                     int state_kind,
                                                     std::shared_ptr<LeafEvaluator> evaluator)
    : BehaviorUCTBase(params, state_kind, evaluator),
      behavior_hypotheses_(h),
      use_true_behaviors_as_hypothesis_(params_->GetBool(
          "BehaviorUctHypothesis::PredictionSettings::UseTrueBehaviorsAsHypothesis", false)),
      belief_tracker_(mcts_parameters_) {}

void BehaviorUCTHypothesisBase::UpdateBeliefs(const ObservedWorld& observed_world) {
  if (use_true_behaviors_as_hypothesis_ || behavior_hypotheses_.empty()) return;
  // belief_update(last_state, new_state): P = last_state.get_probability(h, agent,
  // new_state.get_last_action(agent)) (mcts_state_hypothesis.cpp:116-135); first call: same state
  const ObservedWorld& prev = last_world_ ? *last_world_ : observed_world;
  SearchWorld sw = BuildSearchWorld(prev, behavior_hypotheses_);
  const int A = sw.scenario.n_agents, H = sw.scenario.n_hypotheses;
  std::vector<float> action((size_t)A, 0.0f);
  for (int i = 0; i < A; ++i) {
    const Agent* cur = observed_world.GetAgent(sw.slot_ids[i]);
    action[i] = cur ? (float)cur->last_action.acc : 0.0f;
  }
  bmcts_config cfg = engine_config_;
  int st = evaluator_->Configure(cfg, sw.scenario);
  if (st != BMCTS_OK) throw std::runtime_error("engine configure failed: " + evaluator_->LastError());
  const BehaviorHypothesisIDM& h0 = behavior_hypotheses_[0];
  std::vector<float> prob;
  st = evaluator_->Likelihood(sw.pose, sw.alive, action, h0.num_samples, h0.num_buckets,
                              h0.buckets_lower_bound, h0.buckets_upper_bound,
                              mcts_parameters_.RANDOM_SEED + (uint64_t)(observed_world.world_time * 1000.0),
                              &prob);
  if (st != BMCTS_OK) throw std::runtime_error("likelihood failed: " + evaluator_->LastError());
  std::map<AgentId, std::vector<double>> probs;
  for (int i = 1; i < A; ++i) {
    if (!observed_world.GetAgent(sw.slot_ids[i])) continue;  // agent left the world
    std::vector<double> p(H);
    for (int h = 0; h < H; ++h) p[h] = prob[(size_t)i * H + h];
    probs[sw.slot_ids[i]] = p;
  }
  belief_tracker_.belief_update(probs);
  last_world_ = std::make_shared<ObservedWorld>(observed_world);
  current_beliefs_ = belief_tracker_.get_beliefs();
}

void BehaviorUCTHypothesisBase::DefineTrueBehaviorsAsHypothesis(const ObservedWorld& observed_world) {
  // behavior_uct_hypothesis_base.hpp:80-90; GetOtherAgents() iterates in ascending AgentId
  std::vector<const Agent*> others;
  for (const Agent& a : observed_world.agents)
    if (a.id != observed_world.ego_id) others.push_back(&a);
  std::sort(others.begin(), others.end(), [](const Agent* l, const Agent* r) { return l->id < r->id; });
  const BehaviorHypothesisIDM settings =
      behavior_hypotheses_.empty() ? BehaviorHypothesisIDM() : behavior_hypotheses_[0];
  behavior_hypotheses_.clear();
  std::map<AgentId, unsigned> fixed;
  for (const Agent* a : others) {
    if (!a->has_behavior_model)
      throw std::runtime_error("UseTrueBehaviorsAsHypothesis: agent " + std::to_string(a->id) +
                               " carries no behaviour model");
    if (behavior_hypotheses_.size() >= BMCTS_MAX_HYPOTHESES)
      throw std::runtime_error("UseTrueBehaviorsAsHypothesis: more agents than hypothesis slots");
    fixed[a->id] = (unsigned)behavior_hypotheses_.size();
    BehaviorHypothesisIDM h = settings;  // histogram settings are irrelevant: no belief update
    h.region = a->behavior_model;
    behavior_hypotheses_.push_back(h);
  }
  belief_tracker_.update_fixed_hypothesis_set(fixed);
}

Mcts::HypothesisSampler BehaviorUCTHypothesisBase::MakeSampler(const SearchWorld& sw,
                                                               unsigned tree_index) {
  auto sampler = std::make_shared<HypothesisBeliefTracker::SlotSampler>(
      belief_tracker_.SamplerForSlots(sw.slot_ids, (unsigned)sw.scenario.n_hypotheses, tree_index));
  return [sampler](std::vector<uint8_t>* hyp) { (*sampler)(hyp); };
}

// ----------------------------------------------------------- BehaviorUCTHypothesis

Trajectory BehaviorUCTHypothesis::Plan(double min_planning_time, const ObservedWorld& observed_world) {
  // behavior_uct_hypothesis.cpp:27-106
  ObservedWorld mcts_world = FilterAgents(observed_world);
  last_nearest_agents_.clear();
  for (const Agent& a : mcts_world.agents) last_nearest_agents_.push_back(a.id);
  if (use_true_behaviors_as_hypothesis_) DefineTrueBehaviorsAsHypothesis(observed_world);  // :34-36
  UpdateBeliefs(observed_world);  // all agents, not only the filtered ones (:56)
  SearchWorld sw = BuildSearchWorld(mcts_world, behavior_hypotheses_);
  last_num_actions_ = sw.scenario.n_ego_actions;
  int st = evaluator_->Configure(engine_config_, sw.scenario);
  if (st != BMCTS_OK) throw std::runtime_error("engine configure failed: " + evaluator_->LastError());
  const long trees = std::max<long>(1, mcts_parameters_.NUM_PARALLEL_MCTS);
  RootStats merged = RunTrees(sw, trees, [&](long t, LeafEvaluator* ev) {
    const unsigned tree = tree_index_ * (unsigned)trees + (unsigned)t;
    auto mcts = std::make_unique<Mcts>(mcts_parameters_, EgoMode::kUct, OtherMode::kHypothesis,
                                       sw.scenario.n_agents, sw.scenario.n_ego_actions,
                                       sw.scenario.n_hypotheses, 1, ev, tree);
    mcts->StartSearch(sw.pose, sw.alive, 1, MakeSampler(sw, tree));
    return mcts;
  }, CollectTree);
  DecideFromRootStats(merged);
  FinishPlan(sw, observed_world, last_motion_idx_, min_planning_time);
  return last_trajectory_;
}

// ------------------------------------------------------- BehaviorUCTRiskConstraint

BehaviorUCTRiskConstraint::BehaviorUCTRiskConstraint(const ParamsPtr& params,
                                                     const std::vector<BehaviorHypothesisIDM>& h,
                                                     const ScenarioRiskFunctionPtr& f, int device)
    : BehaviorUCTHypothesisBase(params, h, BMCTS_STATE_RISK_CONSTRAINT, device),
      scenario_risk_function_(f) {
  InitRisk();
}

BehaviorUCTRiskConstraint::BehaviorUCTRiskConstraint(const ParamsPtr& params,
                                                     const std::vector<BehaviorHypothesisIDM>& h,
                                                     const ScenarioRiskFunctionPtr& f,
                                                     std::shared_ptr<LeafEvaluator> evaluator)
    : BehaviorUCTHypothesisBase(params, h, BMCTS_STATE_RISK_CONSTRAINT, evaluator),
      scenario_risk_function_(f) {
  InitRisk();
}

void BehaviorUCTRiskConstraint::InitRisk() {
  // behavior_uct_risk_constraint.cpp:24-35
  default_available_risk_ = params_->GetListFloat("BehaviorUctRiskConstraint::DefaultAvailableRisk", {0.0});
  current_scenario_risk_ = default_available_risk_;
  estimate_scenario_risk_ = params_->GetBool("BehaviorUctRiskConstraint::EstimateScenarioRisk", false);
  update_scenario_risk_ = params_->GetBool("BehaviorUctRiskConstraint::UpdateScenarioRisk", true);
  initialized_available_risk_ = !estimate_scenario_risk_;
}

double BehaviorUCTRiskConstraint::CalculateAvailableScenarioRisk() const {
  // behavior_uct_risk_constraint.cpp:51-72: sum_h mean-belief(h) * integral(h) / area(h)
  const Beliefs& beliefs = belief_tracker_.get_beliefs();
  if (beliefs.empty() || !scenario_risk_function_) return 0.0;
  std::vector<double> belief_sum(beliefs.begin()->second.size(), 0.0);
  for (const auto& kv : beliefs)
    for (size_t h = 0; h < kv.second.size() && h < belief_sum.size(); ++h) belief_sum[h] += kv.second[h];
  double risk = 0.0;
  for (size_t h = 0; h < behavior_hypotheses_.size() && h < belief_sum.size(); ++h) {
    const double integral = scenario_risk_function_->CalculateIntegralValue(behavior_hypotheses_[h]);
    const double area = 1.0 / behavior_hypotheses_[h].ParameterSamplingProbability();
    risk += 1.0 / (double)belief_sum.size() * belief_sum[h] * integral / area;
  }
  return risk;
}

Trajectory BehaviorUCTRiskConstraint::Plan(double min_planning_time, const ObservedWorld& observed_world) {
  // PlanWithMcts (behavior_uct_risk_constraint.hpp:101-217)
  ObservedWorld mcts_world = FilterAgents(observed_world);
  last_nearest_agents_.clear();
  for (const Agent& a : mcts_world.agents) last_nearest_agents_.push_back(a.id);
  if (use_true_behaviors_as_hypothesis_) DefineTrueBehaviorsAsHypothesis(observed_world);  // :107-109
  UpdateBeliefs(observed_world);
  if (estimate_scenario_risk_ && !initialized_available_risk_ && belief_tracker_.beliefs_initialized()) {
    current_scenario_risk_[0] = CalculateAvailableScenarioRisk();  // :136-142
    initialized_available_risk_ = true;
  }
  SearchWorld sw = BuildSearchWorld(mcts_world, behavior_hypotheses_);
  last_num_actions_ = sw.scenario.n_ego_actions;
  int st = evaluator_->Configure(engine_config_, sw.scenario);
  if (st != BMCTS_OK) throw std::runtime_error("engine configure failed: " + evaluator_->LastError());
  const int n_costs = engine_config_.split_safe_dist_collision ? 2 : 1;  // get_num_costs
  MctsParameters current = mcts_parameters_;
  current.cost_constrained_statistic.COST_CONSTRAINTS = current_scenario_risk_;  // :145-146
  const long trees = std::max<long>(1, mcts_parameters_.NUM_PARALLEL_MCTS);
  RootStats merged = RunTrees(sw, trees, [&](long t, LeafEvaluator* ev) {
    const unsigned tree = tree_index_ * (unsigned)trees + (unsigned)t;
    auto mcts = std::make_unique<Mcts>(current, EgoMode::kCostConstrained, OtherMode::kHypothesis,
                                       sw.scenario.n_agents, sw.scenario.n_ego_actions,
                                       sw.scenario.n_hypotheses, n_costs, ev, tree);
    mcts->set_cost_constraints(current_scenario_risk_);
    if (node_prior_)
      mcts->set_node_prior(node_prior_.get(), GetObserver(&sw.scenario),
                           (double)params_->GetInt("BehaviorUctBase::Mcts::NeuralStatistic::PriorVisits", 1));
    mcts->StartSearch(sw.pose, sw.alive, 1, MakeSampler(sw, tree));
    return mcts;
  }, [n_costs](Mcts& mcts) {
    TreeResult r = CollectTree(mcts);
    r.lambdas = mcts.lambdas();
    r.lambdas.resize(n_costs, 0.0);
    return r;
  }, &last_lambdas_);
  DecideFromRootStats(merged);
  FinishPlan(sw, observed_world, last_motion_idx_, min_planning_time);
  return last_trajectory_;
}

void BehaviorUCTRiskConstraint::DecideFromRootStats(const RootStats& merged) {
  last_root_stats_ = merged;
  const int n = (int)merged.count.size();
  const int n_costs = engine_config_.split_safe_dist_collision ? 2 : 1;
  MctsParameters current = mcts_parameters_;
  current.cost_constrained_statistic.COST_CONSTRAINTS = current_scenario_risk_;
  CostConstrainedStatistic root;
  root.Init(n, n_costs, current);
  for (int a = 0; a < n; ++a) {
    root.reward_statistic().MergeRaw(a, merged.count[a], merged.ret_sum[a]);
    root.cost_statistic(0).MergeRaw(a, merged.count[a], merged.cost0_sum[a]);
    if (n_costs > 1) root.cost_statistic(1).MergeRaw(a, merged.count[a], merged.cost1_sum[a]);
  }
  root.set_constraints(current_scenario_risk_);
  std::mt19937 gen(mcts_parameters_.RANDOM_SEED + 17u + (unsigned)merged.iterations);
  std::vector<double> lambdas = last_lambdas_;
  if (lambdas.empty()) lambdas = mcts_parameters_.cost_constrained_statistic.LAMBDAS;
  // greedy_policy(0, ACTION_FILTER_FACTOR, false) (:151-152)
  std::pair<int, Policy> sampled =
      root.GreedyPolicy(0.0, mcts_parameters_.cost_constrained_statistic.ACTION_FILTER_FACTOR, lambdas, gen);
  last_policy_sampled_ = {(unsigned)sampled.first, sampled.second};
  last_expected_risk_ = root.ExpectedPolicyCost(sampled.second);   // :153
  last_return_values_ = root.reward_statistic().GetPolicy();
  last_cost_values_.clear();
  if (n_costs == 2) {  // :163-168
    last_cost_values_["collision"] = root.cost_statistic(1).GetPolicy();
    last_cost_values_["envelope"] = root.cost_statistic(0).GetPolicy();
  } else {
    last_cost_values_["cost"] = root.cost_statistic(0).GetPolicy();
  }
  last_scenario_risk_ = current_scenario_risk_;
  if (initialized_available_risk_ && update_scenario_risk_) {  // :172-181
    current_scenario_risk_ =
        root.UpdatedConstraints(sampled, current_scenario_risk_, mcts_parameters_.DISCOUNT_FACTOR);
    if (current_scenario_risk_.size() > 1 && default_available_risk_.size() > 1)
      current_scenario_risk_[1] = default_available_risk_[1];
    for (double& r : current_scenario_risk_) r = std::min(std::max(0.0, r), 1.0);
  }
  last_motion_idx_ = (unsigned)sampled.first;
  last_num_iterations_ = (unsigned)merged.iterations;
  last_num_nodes_ = (unsigned)merged.nodes;
}

// ---------------------------------------------------------- BehaviorUCTCooperative

Trajectory BehaviorUCTCooperative::Plan(double min_planning_time, const ObservedWorld& observed_world) {
  // behavior_uct_cooperative.cpp:18-77: every agent gets the ego macro-action model
  ObservedWorld mcts_world = FilterAgents(observed_world);
  last_nearest_agents_.clear();
  for (const Agent& a : mcts_world.agents) last_nearest_agents_.push_back(a.id);
  SearchWorld sw = BuildSearchWorld(mcts_world, {});
  last_num_actions_ = sw.scenario.n_ego_actions;
  int st = evaluator_->Configure(engine_config_, sw.scenario);
  if (st != BMCTS_OK) throw std::runtime_error("engine configure failed: " + evaluator_->LastError());
  const long trees = std::max<long>(1, mcts_parameters_.NUM_PARALLEL_MCTS);
  RootStats merged = RunTrees(sw, trees, [&](long t, LeafEvaluator* ev) {
    const unsigned tree = tree_index_ * (unsigned)trees + (unsigned)t;
    auto mcts = std::make_unique<Mcts>(mcts_parameters_, EgoMode::kUct, OtherMode::kUct,
                                       sw.scenario.n_agents, sw.scenario.n_ego_actions, 0, 1, ev, tree);
    mcts->StartSearch(sw.pose, sw.alive, 1, nullptr);
    return mcts;
  }, CollectTree);
  DecideFromRootStats(merged);
  FinishPlan(sw, observed_world, last_motion_idx_, min_planning_time);
  return last_trajectory_;
}

}  // namespace bark_mcts_b200
