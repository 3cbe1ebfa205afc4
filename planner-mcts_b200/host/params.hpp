// params.hpp -- minimal stand-in for BARK's hierarchical ParamServer: "A::B::C" keys,
// typed getters with the default given at the point of use (the way the reference
// reads every key, e.g. behavior_uct_base.cpp:13-39), plus the MctsParameters struct
// the reference fills from it (param_config/mcts_parameters_from_param_server.hpp:30-110).
#pragma once

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "bmcts_c.h"

namespace bark_mcts_b200 {

class Params {
 public:
  void SetReal(const std::string& key, double v) { values_[key] = {v}; }
  void SetInt(const std::string& key, long v) { values_[key] = {(double)v}; }
  void SetBool(const std::string& key, bool v) { values_[key] = {v ? 1.0 : 0.0}; }
  void SetListFloat(const std::string& key, const std::vector<double>& v) { values_[key] = v; }

  double GetReal(const std::string& key, double def) const {
    auto it = values_.find(key);
    return (it == values_.end() || it->second.empty()) ? def : it->second[0];
  }
  long GetInt(const std::string& key, long def) const { return (long)GetReal(key, (double)def); }
  bool GetBool(const std::string& key, bool def) const { return GetReal(key, def ? 1.0 : 0.0) != 0.0; }
  std::vector<double> GetListFloat(const std::string& key, const std::vector<double>& def) const {
    auto it = values_.find(key);
    return it == values_.end() ? def : it->second;
  }
  bool Has(const std::string& key) const { return values_.count(key) != 0; }
  const std::map<std::string, std::vector<double>>& values() const { return values_; }

 private:
  std::map<std::string, std::vector<double>> values_;
};
using ParamsPtr = std::shared_ptr<Params>;

// mcts::MctsParameters (mamcts mcts_parameters.h) as filled by
// MctsParametersFromParamServer; all keys live under "BehaviorUctBase::".
struct MctsParameters {
  double DISCOUNT_FACTOR = 0.9;
  unsigned RANDOM_SEED = 1000;
  long MAX_SEARCH_TIME = 2000;  // ms
  long MAX_NUMBER_OF_ITERATIONS = 2000;
  long MAX_NUMBER_OF_NODES = 2000;
  long MAX_SEARCH_DEPTH = 1000;
  bool USE_BOUND_ESTIMATION = true;
  long NUM_PARALLEL_MCTS = 1;
  bool USE_MULTI_THREADING = false;
  struct {
    long MAX_SEARCH_TIME = 10;
    long MAX_NUMBER_OF_ITERATIONS = 1000;
  } random_heuristic;
  struct {
    double LOWER_BOUND = -1000, UPPER_BOUND = 100, EXPLORATION_CONSTANT = 0.7;
    double PROGRESSIVE_WIDENING_K = 1, PROGRESSIVE_WIDENING_ALPHA = 0.1;
  } uct_statistic;
  struct {
    bool COST_BASED_ACTION_SELECTION = false;
    double LOWER_COST_BOUND = 0, UPPER_COST_BOUND = 1;
    bool PROGRESSIVE_WIDENING_HYPOTHESIS_BASED = true;
    double PROGRESSIVE_WIDENING_ALPHA = 0.25, PROGRESSIVE_WIDENING_K = 4;
    double EXPLORATION_CONSTANT = 0.7;
  } hypothesis_statistic;
  struct {
    unsigned RANDOM_SEED_HYPOTHESIS_SAMPLING = 2000;
    unsigned HISTORY_LENGTH = 10;
    double PROBABILITY_DISCOUNT = 0.7;
    int POSTERIOR_TYPE = 1;  // 0 product, 1 sum
  } hypothesis_belief_tracker;
  struct {
    std::vector<double> LAMBDAS = {1.0};
    double COST_UPPER_BOUND = 1, COST_LOWER_BOUND = 0;
    double REWARD_UPPER_BOUND = 100, REWARD_LOWER_BOUND = -1000;
    std::vector<double> COST_CONSTRAINTS = {0.0};
    double KAPPA = 10.0, GRADIENT_UPDATE_STEP = 1.0, TAU_GRADIENT_CLIP = 1.0;
    double ACTION_FILTER_FACTOR = 0.5;
    double EXPLORATION_REDUCTION_FACTOR = 0.5, EXPLORATION_REDUCTION_CONSTANT_OFFSET = 20.0,
           EXPLORATION_REDUCTION_INIT = 200.0;
    long MIN_VISITS_POLICY_READY = -1;
    std::vector<bool> USE_COST_THRESHOLDING = {false, false};
    std::vector<bool> USE_CHANCE_CONSTRAINED_UPDATES = {false, false};
    std::vector<double> COST_THRESHOLDS = {0.1, 0.0};
    bool USE_LAMBDA_POLICY = true;
    long MAX_SOLVER_TIME = -1;
  } cost_constrained_statistic;
  // engine-side knob (not a reference key): leaves evaluated per GPU launch.  1 = the
  // reference's strictly sequential search.
  long WAVE_SIZE = 0;  // 0 = automatic: MaxNumIterations / 128, between 1 and 256
  // engine-side knob: host threads of a multi-threaded root-parallel search (0 = all cores);
  // with more trees than threads each thread interleaves several trees
  long MAX_THREADS = 0;
};

// param_config/mcts_parameters_from_param_server.hpp:30-110 (keys relative to `prefix`)
inline MctsParameters MctsParametersFromParamServer(const Params& p, const std::string& prefix) {
  MctsParameters m;
  auto k = [&](const char* s) { return prefix + s; };
  m.DISCOUNT_FACTOR = p.GetReal(k("Mcts::DiscountFactor"), 0.9);
  m.RANDOM_SEED = (unsigned)p.GetInt(k("Mcts::RandomSeed"), 1000);
  m.MAX_SEARCH_TIME = p.GetInt(k("Mcts::MaxSearchTime"), 2000);
  m.MAX_NUMBER_OF_ITERATIONS = p.GetInt(k("Mcts::MaxNumIterations"), 2000);
  m.MAX_NUMBER_OF_NODES = p.GetInt(k("Mcts::MaxNumNodes"), 2000);
  m.MAX_SEARCH_DEPTH = p.GetInt(k("Mcts::MaxSearchDepth"), 1000);
  m.USE_BOUND_ESTIMATION = p.GetBool(k("Mcts::UseBoundEstimation"), true);
  m.NUM_PARALLEL_MCTS = p.GetInt(k("Mcts::NumParallelMcts"), 1);
  m.USE_MULTI_THREADING = p.GetBool(k("Mcts::UseMultiThreading"), false);
  m.random_heuristic.MAX_SEARCH_TIME = p.GetInt(k("Mcts::RandomHeuristic::MaxSearchTime"), 10);
  m.random_heuristic.MAX_NUMBER_OF_ITERATIONS =
      p.GetInt(k("Mcts::RandomHeuristic::MaxNumIterations"), 1000);
  m.uct_statistic.LOWER_BOUND = p.GetReal(k("Mcts::ReturnLowerBound"), -1000);
  m.uct_statistic.UPPER_BOUND = p.GetReal(k("Mcts::ReturnUpperBound"), 100);
  m.uct_statistic.EXPLORATION_CONSTANT = p.GetReal(k("Mcts::UctStatistic::ExplorationConstant"), 0.7);
  m.uct_statistic.PROGRESSIVE_WIDENING_K =
      p.GetReal(k("Mcts::UctStatistic::ProgressiveWidening::K"), 1);
  m.uct_statistic.PROGRESSIVE_WIDENING_ALPHA =
      p.GetReal(k("Mcts::UctStatistic::ProgressiveWidening::Alpha"), 0.1);
  m.hypothesis_statistic.COST_BASED_ACTION_SELECTION =
      p.GetBool(k("Mcts::HypothesisStatistic::CostBasedActionSelection"), false);
  m.hypothesis_statistic.LOWER_COST_BOUND = p.GetReal(k("Mcts::LowerCostBound"), 0);
  m.hypothesis_statistic.UPPER_COST_BOUND = p.GetReal(k("Mcts::UpperCostBound"), 1);
  m.hypothesis_statistic.PROGRESSIVE_WIDENING_HYPOTHESIS_BASED =
      p.GetBool(k("Mcts::HypothesisStatistic::ProgressiveWidening::HypothesisBased"), true);
  m.hypothesis_statistic.PROGRESSIVE_WIDENING_ALPHA =
      p.GetReal(k("Mcts::HypothesisStatistic::ProgressiveWidening::Alpha"), 0.25);
  m.hypothesis_statistic.PROGRESSIVE_WIDENING_K =
      p.GetReal(k("Mcts::HypothesisStatistic::ProgressiveWidening::K"), 4);
  m.hypothesis_statistic.EXPLORATION_CONSTANT =
      p.GetReal(k("Mcts::HypothesisStatistic::ExplorationConstant"), 0.7);
  // BeliefTrackerParametersFromParamServer (:16-27)
  m.hypothesis_belief_tracker.RANDOM_SEED_HYPOTHESIS_SAMPLING =
      (unsigned)p.GetInt(k("Mcts::BeliefTracker::RandomSeedHypSampling"), 2000);
  m.hypothesis_belief_tracker.HISTORY_LENGTH =
      (unsigned)p.GetInt(k("Mcts::BeliefTracker::HistoryLength"), 10);
  m.hypothesis_belief_tracker.PROBABILITY_DISCOUNT =
      p.GetReal(k("Mcts::BeliefTracker::ProbabilityDiscount"), 0.7);
  m.hypothesis_belief_tracker.POSTERIOR_TYPE =
      (int)p.GetInt(k("Mcts::BeliefTracker::PosteriorType"), 1);
  auto& c = m.cost_constrained_statistic;
  c.LAMBDAS = p.GetListFloat(k("Mcts::CostConstrainedStatistic::LambdaInit"), {1.0});
  c.COST_UPPER_BOUND = m.hypothesis_statistic.UPPER_COST_BOUND;
  c.COST_LOWER_BOUND = m.hypothesis_statistic.LOWER_COST_BOUND;
  c.REWARD_UPPER_BOUND = m.uct_statistic.UPPER_BOUND;
  c.REWARD_LOWER_BOUND = m.uct_statistic.LOWER_BOUND;
  c.COST_CONSTRAINTS = {0.0};  // set per Plan()
  c.KAPPA = p.GetReal(k("Mcts::CostConstrainedStatistic::Kappa"), 10.0);
  c.GRADIENT_UPDATE_STEP = p.GetReal(k("Mcts::CostConstrainedStatistic::GradientUpdateScaling"), 1.0);
  c.TAU_GRADIENT_CLIP = p.GetReal(k("Mcts::CostConstrainedStatistic::TauGradientClip"), 1.0);
  c.ACTION_FILTER_FACTOR = p.GetReal(k("Mcts::CostConstrainedStatistic::ActionFilterFactor"), 0.5);
  c.EXPLORATION_REDUCTION_FACTOR =
      p.GetReal(k("Mcts::CostConstrainedStatistic::ExplorationReductionFactor"), 0.5);
  c.EXPLORATION_REDUCTION_CONSTANT_OFFSET =
      p.GetReal(k("Mcts::CostConstrainedStatistic::ExplorationReductionOffset"), 20.0);
  c.EXPLORATION_REDUCTION_INIT =
      p.GetReal(k("Mcts::CostConstrainedStatistic::ExplorationReductionInit"), 200.0);
  c.MIN_VISITS_POLICY_READY = p.GetInt(k("Mcts::CostConstrainedStatistic::MinVisitsPolicyReady"), -1);
  auto to_bool = [](const std::vector<double>& v) {
    std::vector<bool> b;
    for (double x : v) b.push_back(x == 1.0);
    return b;
  };
  c.USE_COST_THRESHOLDING =
      to_bool(p.GetListFloat(k("Mcts::CostConstrainedStatistic::UseCostTresholding"), {0.0, 0.0}));
  c.USE_CHANCE_CONSTRAINED_UPDATES = to_bool(
      p.GetListFloat(k("Mcts::CostConstrainedStatistic::UseChanceConstrainedUpdate"), {0.0, 0.0}));
  c.COST_THRESHOLDS = p.GetListFloat(k("Mcts::CostConstrainedStatistic::CostThresholds"), {0.1, 0.0});
  c.USE_LAMBDA_POLICY = p.GetBool(k("Mcts::CostConstrainedStatistic::UseLambdaPolicy"), true);
  c.MAX_SOLVER_TIME = p.GetInt(k("Mcts::CostConstrainedStatistic::MaxSolverTime"), -1);
  // iterations selected per engine batch.  Root values are means under the tree policy, and a
  // wave selects without seeing the results of its own earlier selections: waves that are
  // small against the iteration budget (<= 1 %) reproduce the sequential search's root values
  // within Monte-Carlo noise (tests/test_planner.py), larger ones trade bias for throughput.
  m.WAVE_SIZE = p.GetInt(k("Mcts::Engine::WaveSize"), 0);
  m.MAX_THREADS = p.GetInt(k("Mcts::Engine::MaxThreads"), 0);
  return m;
}

// MctsStateParametersFromParamServer + MctsEvaluationParametersFromParamServer
// (param_config/mcts_state_parameters_from_param_server.hpp:19-51), the evaluator subtree
// (tests/data/nheuristic_test.json:281-311) and the ego model (:152-199) -> bmcts_config
inline bmcts_config EngineConfigFromParamServer(const Params& p, const std::string& prefix,
                                                int state_kind, const MctsParameters& m) {
  bmcts_config c;
  bmcts_config_default(&c);
  auto k = [&](const char* s) { return prefix + s; };
  c.goal_reward = (float)p.GetReal(k("Mcts::State::GoalReward"), 100.0);
  c.collision_reward = (float)p.GetReal(k("Mcts::State::CollisionReward"), -1000.0);
  c.safe_dist_violated_reward = (float)p.GetReal(k("Mcts::State::SafeDistViolatedReward"), -0.1);
  c.out_of_drivable_reward = (float)p.GetReal(k("Mcts::State::DrivableCollisionReward"), 0.0);
  c.goal_cost = (float)p.GetReal(k("Mcts::State::GoalCost"), -100.0);
  c.collision_cost = (float)p.GetReal(k("Mcts::State::CollisionCost"), 1000.0);
  c.safe_dist_violated_cost = (float)p.GetReal(k("Mcts::State::SafeDistViolatedCost"), 0.1);
  c.out_of_drivable_cost = (float)p.GetReal(k("Mcts::State::DrivableCollisionCost"), 0.0);
  c.cooperation_factor = (float)p.GetReal(k("Mcts::State::CooperationFactor"), 0.2);
  c.step_reward = (float)p.GetReal(k("Mcts::State::StepReward"), 0.0);
  c.prediction_k = (float)p.GetReal(k("Mcts::State::PredictionK"), 0.5);
  c.prediction_alpha = (float)p.GetReal(k("Mcts::State::PredictionAlpha"), 0.0);
  c.normalization_tau = (float)p.GetReal(k("Mcts::State::NormalizationTau"), 0.2);
  c.split_safe_dist_collision = p.GetBool(k("Mcts::State::SplitSafeDistCollision"), false);
  c.chance_costs = p.GetBool(k("Mcts::State::ChanceCosts"), false);
  const std::string ev = prefix + "Mcts::State::EvaluationParameters::";
  c.add_safe_dist = p.GetBool(ev + "AddSafeDist", false);
  c.static_safe_dist_is_terminal = p.GetBool(ev + "StaticSafeDistIsTerminal", false);
  c.dynamic_safe_dist_is_terminal = p.GetBool(ev + "DynamicSafeDistIsTerminal", false);
  c.out_of_drivable_is_terminal = p.GetBool(ev + "OutOfDrivableIsTerminal", false);
  const std::string ep = prefix + "Mcts::State::EvaluatorParams::";
  c.drivable_lat_safe_dist = (float)p.GetReal(ep + "EvaluatorSafeDistDrivableArea::LateralSafeDist", 0.5);
  c.drivable_lon_safe_dist =
      (float)p.GetReal(ep + "EvaluatorSafeDistDrivableArea::LongitudinalSafeDist", 0.5);
  c.static_lat_safe_dist = (float)p.GetReal(ep + "EvaluatorStaticSafeDist::LateralSafeDist", 1.5);
  c.static_lon_safe_dist = (float)p.GetReal(ep + "EvaluatorStaticSafeDist::LongitudinalSafeDist", 1.5);
  c.dyn_max_other_decel = (float)p.GetReal(ep + "EvaluatorDynamicSafeDist::MaxOtherDecceleration", 5.0);
  c.dyn_max_ego_decel = (float)p.GetReal(ep + "EvaluatorDynamicSafeDist::MaxEgoDecceleration", 5.0);
  c.dyn_reaction_time_others = (float)p.GetReal(ep + "EvaluatorDynamicSafeDist::ReactionTimeOthers", 0.8);
  c.dyn_reaction_time_ego = (float)p.GetReal(ep + "EvaluatorDynamicSafeDist::ReactionTimeEgo", 0.2);
  c.dyn_to_rear = p.GetBool(ep + "EvaluatorDynamicSafeDist::ToRear", true);
  c.discount_factor = (float)m.DISCOUNT_FACTOR;
  c.max_rollout_steps = (int)m.random_heuristic.MAX_NUMBER_OF_ITERATIONS;
  const std::string eg = prefix + "EgoBehavior::";
  c.wheel_base = (float)p.GetReal(eg + "DynamicModel::wheel_base", 2.7);
  c.delta_max = (float)p.GetReal(eg + "DynamicModel::delta_max", 0.2);
  c.crosstrack_gain = (float)p.GetReal(eg + "BehaviorIDMLaneTracking::CrosstrackErrorGain", 1.0);
  c.lon_acc_min = (float)p.GetReal(eg + "DynamicModel::LonAccelerationMin", -8.0);
  c.lon_acc_max = (float)p.GetReal(eg + "DynamicModel::LonAccelerationMax", 4.0);
  c.ego_max_velocity = (float)p.GetReal(eg + "BehaviorIDMClassic::MaxVelocity", 50.0);
  c.num_substeps = (int)p.GetInt(eg + "BehaviorIDMClassic::NumTrajectoryTimePoints", 11) - 1;
  c.limit_steering_rate = p.GetBool(eg + "BehaviorIDMLaneTracking::LimitSteeringRate", true);
  c.lat_acc_max = (float)p.GetReal(eg + "DynamicModel::LatAccMax", 4.0);
  c.lat_acc_min = (float)p.GetReal(eg + "DynamicModel::LatAccMin", -4.0);
  c.ego_integration_dt =
      (float)p.GetReal(eg + "BehaviorMotionPrimitives::IntegrationTimeDelta", 0.02);
  c.state_kind = state_kind;
  return c;
}

}  // namespace bark_mcts_b200
