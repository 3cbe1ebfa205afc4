// behavior_uct.hpp -- the drop-in planner surface: BehaviorUCTHypothesis,
// BehaviorUCTRiskConstraint, BehaviorUCTCooperative with Plan(delta_time, observed_world),
// Clone() and the reference's parameter keys and debug-info getters
// (behavior_uct_base.hpp:29-55,71; behavior_uct_hypothesis.hpp:20-35;
//  behavior_uct_risk_constraint.hpp:34-92; behavior_uct_cooperative.hpp:19-29),
// over the POD ObservedWorld of world.hpp.  The tree search runs on the host
// (mcts.hpp); leaf evaluation and expansion steps run on the GPU through the C ABI.
#pragma once

#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "evaluator.hpp"
#include "mcts.hpp"
#include "node_prior.hpp"
#include "params.hpp"
#include "statistics.hpp"
#include "world.hpp"

namespace bark_mcts_b200 {

struct UctBaseDebugInfos {  // behavior_uct_base.hpp:29-55
  Policy GetLastReturnValues() const { return last_return_values_; }
  std::vector<AgentId> GetLastNearestAgents() const { return last_nearest_agents_; }
  unsigned GetLastNumIterations() const { return last_num_iterations_; }
  unsigned GetLastNumNodes() const { return last_num_nodes_; }
  double GetLastSolutionTime() const { return last_solution_time_; }
  // tree shape (what the reference's tests derive from GetLastMctsEdgeInfo)
  int GetLastMaxTreeDepth() const { return last_max_tree_depth_; }
  std::vector<int> GetLastRootExpandedActions() const { return last_root_expanded_actions_; }
  // behavior_uct_base.hpp:32-37 (BarkMctsEdgeInfo / BarkMctsStateInfo of the last tree; filled
  // when ExtractEdgeInfo / ExtractStateInfo are set, down to MaxExtractionDepth)
  struct MctsEdgeInfo {
    AgentId agent_id = 0;
    unsigned depth = 0;
    uint32_t action = 0;   // ego / cooperative: macro-action index; hypothesis agents: acceleration bits
    double action_weight = 0.0;
  };
  struct MctsStateInfo {
    unsigned depth = 0;
    std::vector<AgentId> agent_ids;
    std::vector<State> states;
  };
  std::vector<MctsEdgeInfo> GetLastMctsEdgeInfo() const { return mcts_edge_infos_; }
  std::vector<MctsStateInfo> GetLastMctsStateInfo() const { return mcts_state_infos_; }
  std::vector<MctsEdgeInfo> mcts_edge_infos_;
  std::vector<MctsStateInfo> mcts_state_infos_;
  Policy last_return_values_;
  std::vector<AgentId> last_nearest_agents_;
  unsigned last_num_iterations_ = 0, last_num_nodes_ = 0;
  double last_solution_time_ = 0.0;
  int last_max_tree_depth_ = 0;
  std::vector<int> last_root_expanded_actions_;
};

struct UctHypothesisDebugInfos {  // behavior_uct_hypothesis_base.hpp:21-25
  Beliefs GetCurrentBeliefs() const { return current_beliefs_; }
  Beliefs current_beliefs_;
};

using PolicySampled = std::pair<unsigned, Policy>;

struct UctRiskConstraintDebugInfos {  // behavior_uct_risk_constraint.hpp:34-51
  std::vector<double> GetLastScenarioRisk() const { return last_scenario_risk_; }
  std::vector<double> GetLastExpectedRisk() const { return last_expected_risk_; }
  PolicySampled GetLastPolicySampled() const { return last_policy_sampled_; }
  std::map<std::string, Policy> GetLastCostValues() const { return last_cost_values_; }
  PolicySampled last_policy_sampled_;
  std::vector<double> last_expected_risk_, last_scenario_risk_;
  std::map<std::string, Policy> last_cost_values_;
};

// risk_calculation::ScenarioRiskFunction reduced to what the planner calls
// (behavior_uct_risk_constraint.cpp:51-72); the risk-function tooling itself is out of scope
struct ScenarioRiskFunction {
  virtual ~ScenarioRiskFunction() {}
  virtual double CalculateIntegralValue(const BehaviorHypothesisIDM& hypothesis) const = 0;
};
using ScenarioRiskFunctionPtr = std::shared_ptr<ScenarioRiskFunction>;

class BehaviorUCTBase : public UctBaseDebugInfos {
 public:
  BehaviorUCTBase(const ParamsPtr& params, int state_kind, int device);
  BehaviorUCTBase(const ParamsPtr& params, int state_kind, std::shared_ptr<LeafEvaluator> evaluator);
  virtual ~BehaviorUCTBase() {}

  virtual Trajectory Plan(double min_planning_time, const ObservedWorld& observed_world) = 0;

  unsigned GetLastMacroAction() const { return last_motion_idx_; }
  Action GetLastAction() const { return last_action_; }
  Trajectory GetLastTrajectory() const { return last_trajectory_; }
  ParamsPtr GetParams() const { return params_; }
  std::string GetPrimitiveName(unsigned action) const;
  // behavior_uct_base.cpp:54-64: ego + the MaxNearestAgents nearest agents
  ObservedWorld FilterAgents(const ObservedWorld& observed_world) const;
  // statistics of the last search at the root, mergeable over root-parallel ranks
  RootStats GetLastRootStats() const { return last_root_stats_; }
  // re-decide from externally merged statistics (NCCL all-reduce over ranks)
  virtual void DecideFromRootStats(const RootStats& merged);
  // after DecideFromRootStats on merged statistics: the trajectory / action of the (possibly
  // different) chosen macro action, from the world of the last Plan()
  void RefreshTrajectory() {
    if (has_last_plan_) FinishPlan(last_sw_, last_plan_world_, last_motion_idx_, last_plan_time_);
  }
  uint64_t EngineLaunchCount() const {
    uint64_t n = evaluator_ ? evaluator_->LaunchCount() : 0;
    for (const auto& ev : tree_evaluators_) n += ev->LaunchCount();
    return n;
  }
  const MctsParameters& mcts_parameters() const { return mcts_parameters_; }
  void set_tree_index(unsigned i) { tree_index_ = i; }
  // how further rollout engines are made for Mcts::UseMultiThreading (one engine, i.e. one
  // CUDA stream, per tree thread); without a factory the trees run one after the other
  void set_evaluator_factory(std::function<std::shared_ptr<LeafEvaluator>()> f) {
    evaluator_factory_ = [f](int) { return f(); };
  }
  // ... with the index of the engine (engines are numbered thread by thread): the product factory
  // spreads them over the GPUs of Mcts::Engine::Devices, so that ONE Plan() in ONE process uses
  // several GPUs; the root statistics of all trees are merged in this process, no collective
  void set_indexed_evaluator_factory(std::function<std::shared_ptr<LeafEvaluator>(int)> f) {
    evaluator_factory_ = std::move(f);
  }
  // Mcts::Engine::Devices (this engine's key): CUDA device ordinals of the rollout engines of a
  // root-parallel search; empty = the device the planner was constructed with
  std::vector<int> EngineDevices(int default_device) const {
    std::vector<int> d;
    for (double x : params_->GetListFloat("BehaviorUctBase::Mcts::Engine::Devices", {})) d.push_back((int)x);
    if (d.empty()) d.push_back(default_device);
    return d;
  }

 protected:
  struct SearchWorld {
    bmcts_scenario scenario;
    std::vector<AgentId> slot_ids;  // slot 0 = ego, then ascending AgentId
    std::vector<float> pose;        // [A][4]
    uint32_t alive = 0;
  };
  // scenario for a world: map + shapes + hypotheses + the ego's valid macro actions
  SearchWorld BuildSearchWorld(const ObservedWorld& world,
                               const std::vector<BehaviorHypothesisIDM>& hypotheses) const;
  // the chosen macro action executed from the current state (ego_behavior_model_->Plan)
  Trajectory MacroActionTrajectory(const SearchWorld& sw, unsigned action, double delta_time,
                                   double t0, Action* last_action) const;
  void FinishPlan(const SearchWorld& sw, const ObservedWorld& observed_world, unsigned best_action,
                  double min_planning_time);
  // result of one root-parallel tree
  struct TreeResult {
    RootStats stats;
    double time_ms = 0.0;
    int max_depth = 0;
    std::vector<int> root_expanded;
    std::vector<double> lambdas;
    std::vector<Mcts::EdgeInfo> edges;
    std::vector<Mcts::StateInfo> states;
  };
  // runs `trees` independent searches (Mcts::NumParallelMcts) -- on threads with one engine
  // each when Mcts::UseMultiThreading is set, and with the waves of several trees in flight per
  // thread when there are more trees than threads -- and merges their root statistics in tree
  // order.  start_tree builds the search of tree t on the given engine and calls StartSearch.
  using StartTree = std::function<std::unique_ptr<Mcts>(long, LeafEvaluator*)>;
  static constexpr long kMaxTreesInFlight = 8;
  static TreeResult CollectTree(Mcts& mcts);
  RootStats RunTrees(const SearchWorld& sw, long trees, const StartTree& start_tree,
                     const std::function<TreeResult(Mcts&)>& collect,
                     std::vector<double>* mean_lambdas = nullptr);

  ParamsPtr params_;
  int state_kind_;
  SearchWorld last_sw_;             // inputs of the last FinishPlan (RefreshTrajectory)
  ObservedWorld last_plan_world_;
  double last_plan_time_ = 0.0;
  bool has_last_plan_ = false;
  std::shared_ptr<LeafEvaluator> evaluator_;
  std::function<std::shared_ptr<LeafEvaluator>(int)> evaluator_factory_;
  std::vector<std::shared_ptr<LeafEvaluator>> tree_evaluators_;  // engines of the tree threads
  MctsParameters mcts_parameters_;
  bmcts_config engine_config_;
  unsigned max_nearest_agents_;
  int constant_action_idx_;
  bool extract_edge_info_ = true, extract_state_info_ = false;  // behavior_uct_base.cpp:19-39
  unsigned max_extraction_depth_ = 10;
  unsigned last_motion_idx_ = 0;
  unsigned tree_index_ = 0;
  Action last_action_;
  Trajectory last_trajectory_;
  RootStats last_root_stats_;
  int last_num_actions_ = 0;
};

// BehaviorUCTHypothesisBase<T> (behavior_uct_hypothesis_base.hpp:27-114)
class BehaviorUCTHypothesisBase : public BehaviorUCTBase, public UctHypothesisDebugInfos {
 public:
  BehaviorUCTHypothesisBase(const ParamsPtr& params, const std::vector<BehaviorHypothesisIDM>& h,
                            int state_kind, int device);
  BehaviorUCTHypothesisBase(const ParamsPtr& params, const std::vector<BehaviorHypothesisIDM>& h,
                            int state_kind, std::shared_ptr<LeafEvaluator> evaluator);
  const HypothesisBeliefTracker& GetBeliefTracker() const { return belief_tracker_; }
  std::vector<BehaviorHypothesisIDM> GetBehaviorHypotheses() const { return behavior_hypotheses_; }
  // behavior_uct_hypothesis_base.hpp:92-114 (all agents, not only the filtered ones)
  void UpdateBeliefs(const ObservedWorld& observed_world);
  // behavior_uct_hypothesis_base.hpp:80-90: the others' true behaviour models become the
  // hypothesis set, one per agent, fixed (no belief tracking)
  void DefineTrueBehaviorsAsHypothesis(const ObservedWorld& observed_world);
  // forget the observation history (BeliefCalculator::Reset)
  void ResetBeliefs() {
    belief_tracker_ = HypothesisBeliefTracker(mcts_parameters_);
    last_world_.reset();
    current_beliefs_.clear();
  }

 protected:
  Mcts::HypothesisSampler MakeSampler(const SearchWorld& sw, unsigned tree_index);
  std::vector<BehaviorHypothesisIDM> behavior_hypotheses_;
  bool use_true_behaviors_as_hypothesis_;
  HypothesisBeliefTracker belief_tracker_;
  std::shared_ptr<ObservedWorld> last_world_;
};

class BehaviorUCTHypothesis : public BehaviorUCTHypothesisBase {
 public:
  BehaviorUCTHypothesis(const ParamsPtr& params, const std::vector<BehaviorHypothesisIDM>& h,
                        int device = 0)
      : BehaviorUCTHypothesisBase(params, h, BMCTS_STATE_HYPOTHESIS, device) {}
  BehaviorUCTHypothesis(const ParamsPtr& params, const std::vector<BehaviorHypothesisIDM>& h,
                        std::shared_ptr<LeafEvaluator> evaluator)
      : BehaviorUCTHypothesisBase(params, h, BMCTS_STATE_HYPOTHESIS, evaluator) {}
  Trajectory Plan(double min_planning_time, const ObservedWorld& observed_world) override;
  std::shared_ptr<BehaviorUCTBase> Clone() const {
    return std::make_shared<BehaviorUCTHypothesis>(*this);
  }
};

class BehaviorUCTRiskConstraint : public BehaviorUCTHypothesisBase,
                                  public UctRiskConstraintDebugInfos {
 public:
  BehaviorUCTRiskConstraint(const ParamsPtr& params, const std::vector<BehaviorHypothesisIDM>& h,
                            const ScenarioRiskFunctionPtr& scenario_risk_function, int device = 0);
  BehaviorUCTRiskConstraint(const ParamsPtr& params, const std::vector<BehaviorHypothesisIDM>& h,
                            const ScenarioRiskFunctionPtr& scenario_risk_function,
                            std::shared_ptr<LeafEvaluator> evaluator);
  Trajectory Plan(double min_planning_time, const ObservedWorld& observed_world) override;
  void DecideFromRootStats(const RootStats& merged) override;
  std::shared_ptr<BehaviorUCTBase> Clone() const {
    return std::make_shared<BehaviorUCTRiskConstraint>(*this);
  }
  ScenarioRiskFunctionPtr GetScenarioRiskFunction() const { return scenario_risk_function_; }
  double CalculateAvailableScenarioRisk() const;  // behavior_uct_risk_constraint.cpp:51-72
  // neural warm start of the ego statistic (node_prior.hpp); null = plain CostConstrainedStatistic
  void set_node_prior(std::shared_ptr<NodePriorProvider> provider) { node_prior_ = std::move(provider); }
  std::shared_ptr<NodePriorProvider> GetNodePrior() const { return node_prior_; }
  NearestObserver GetObserver(const MapGeometry* map) const {
    NearestObserver o = NearestObserver::FromParams(*params_);
    if (map) o.SetWorldBounds(*map);
    return o;
  }

 protected:
  void InitRisk();
  std::vector<double> default_available_risk_, current_scenario_risk_;
  bool estimate_scenario_risk_, initialized_available_risk_, update_scenario_risk_;
  ScenarioRiskFunctionPtr scenario_risk_function_;
  std::vector<double> last_lambdas_;
  std::shared_ptr<NodePriorProvider> node_prior_;
};

// behavior_uct_nheuristic_risk_constraint.hpp:23-60: PlanWithMcts with
// mcts::NeuralCostConstrainedStatistic as the ego statistic.  The reference's (model file,
// bark-ml observer, value converter) triple is one NodePriorProvider here.
class BehaviorUCTNHeuristicRiskConstraint : public BehaviorUCTRiskConstraint {
 public:
  BehaviorUCTNHeuristicRiskConstraint(const ParamsPtr& params,
                                      const std::vector<BehaviorHypothesisIDM>& h,
                                      const ScenarioRiskFunctionPtr& scenario_risk_function,
                                      std::shared_ptr<NodePriorProvider> node_prior, int device = 0)
      : BehaviorUCTRiskConstraint(params, h, scenario_risk_function, device) {
    set_node_prior(std::move(node_prior));
  }
  BehaviorUCTNHeuristicRiskConstraint(const ParamsPtr& params,
                                      const std::vector<BehaviorHypothesisIDM>& h,
                                      const ScenarioRiskFunctionPtr& scenario_risk_function,
                                      std::shared_ptr<NodePriorProvider> node_prior,
                                      std::shared_ptr<LeafEvaluator> evaluator)
      : BehaviorUCTRiskConstraint(params, h, scenario_risk_function, evaluator) {
    set_node_prior(std::move(node_prior));
  }
  Trajectory Plan(double min_planning_time, const ObservedWorld& observed_world) override {
    if (!node_prior_) throw std::runtime_error("BehaviorUCTNHeuristicRiskConstraint: no node prior set");
    return BehaviorUCTRiskConstraint::Plan(min_planning_time, observed_world);
  }
  std::shared_ptr<BehaviorUCTBase> Clone() const {
    return std::make_shared<BehaviorUCTNHeuristicRiskConstraint>(*this);
  }
};

// belief_calculator/belief_calculator.hpp:23-47: the beliefs over a hypothesis set from a sequence
// of observed worlds, without a planner around them (the likelihoods come from the same engine
// call as in BehaviorUCTHypothesisBase::UpdateBeliefs)
class BeliefCalculator {
 public:
  BeliefCalculator(const ParamsPtr& params, const std::vector<BehaviorHypothesisIDM>& h, int device = 0)
      : impl_(params, h, device) {}
  BeliefCalculator(const ParamsPtr& params, const std::vector<BehaviorHypothesisIDM>& h,
                   std::shared_ptr<LeafEvaluator> evaluator)
      : impl_(params, h, evaluator) {}
  void BeliefUpdate(const ObservedWorld& observed_world) { impl_.UpdateBeliefs(observed_world); }
  Beliefs GetBeliefs() const { return impl_.GetCurrentBeliefs(); }
  std::vector<BehaviorHypothesisIDM> GetBehaviorHypotheses() const { return impl_.GetBehaviorHypotheses(); }
  ParamsPtr GetParams() const { return impl_.GetParams(); }
  void Reset() { impl_.ResetBeliefs(); }

 private:
  BehaviorUCTHypothesis impl_;
};

class BehaviorUCTCooperative : public BehaviorUCTBase {
 public:
  explicit BehaviorUCTCooperative(const ParamsPtr& params, int device = 0)
      : BehaviorUCTBase(params, BMCTS_STATE_COOPERATIVE, device) {}
  BehaviorUCTCooperative(const ParamsPtr& params, std::shared_ptr<LeafEvaluator> evaluator)
      : BehaviorUCTBase(params, BMCTS_STATE_COOPERATIVE, evaluator) {}
  Trajectory Plan(double min_planning_time, const ObservedWorld& observed_world) override;
  std::shared_ptr<BehaviorUCTBase> Clone() const {
    return std::make_shared<BehaviorUCTCooperative>(*this);
  }
};

}  // namespace bark_mcts_b200
