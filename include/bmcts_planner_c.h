/*
 * bmcts_planner_c.h -- C ABI of the BehaviorUCT* planners (host tree search + GPU leaf
 * evaluation) for bindings that cannot use the C++ classes of
 * planner-mcts_b200/host/behavior_uct.hpp directly (Python ctypes in this repo's tests and
 * bench).  It mirrors the reference's pybind surface
 * (/root/reference/bark_mcts/python_wrapper/python_planner_uct.cpp:131-248): construct a
 * planner from parameters (+ hypotheses), call Plan(delta_time, observed_world), read the
 * last_* debug infos.
 */
#ifndef BMCTS_PLANNER_C_H_
#define BMCTS_PLANNER_C_H_

#include "bmcts_c.h"

#ifdef __cplusplus
extern "C" {
#endif

enum {
  BMCTS_PLANNER_HYPOTHESIS = 0,      /* BehaviorUCTHypothesis */
  BMCTS_PLANNER_RISK_CONSTRAINT = 1, /* BehaviorUCTRiskConstraint */
  BMCTS_PLANNER_COOPERATIVE = 2,     /* BehaviorUCTCooperative */
  /* BehaviorUCTNHeuristicRiskConstraint (behavior_uct_nheuristic_risk_constraint.hpp:23-60):
   * the risk-constrained planner whose ego nodes are warm-started by a learned model; needs
   * bmcts_planner_set_node_prior before the first plan */
  BMCTS_PLANNER_NHEURISTIC_RISK_CONSTRAINT = 3
};

/* what a node-prior callback filled */
enum { BMCTS_PRIOR_POLICY = 1, BMCTS_PRIOR_VALUES = 2 };

/* The learned model behind NeuralCostConstrainedStatistic::initialize_from_neural_model
 * (mcts_statistics/mcts_neural_cost_constrained_statistic.hpp:175-198), batched: called once
 * per wave of the search with the observations [n_nodes][obs_dim] of every node the wave
 * created (bmcts_planner_observe layout).  Fill, per node and ego action ([n_nodes][n_actions]):
 * the exploration policy and / or the value triple {return, envelope risk, collision risk};
 * return BMCTS_PRIOR_POLICY | BMCTS_PRIOR_VALUES for what was filled, < 0 on failure.  Called
 * from the search threads, one call at a time. */
typedef int (*bmcts_node_prior_fn)(void* user, int n_nodes, int obs_dim, const float* observations,
                                   int n_actions, float* policy, float* returns,
                                   float* envelope_risk, float* collision_risk);

#define BMCTS_PLAN_MAX_TRAJ 64

/* one agent of the observed world (BARK Agent: state, shape, last action) */
typedef struct bmcts_agent_state {
  uint32_t id;
  float x, y, theta, v;
  float length, width, rear;
  float last_acc; /* GetStateInputHistory().back().second, longitudinal part */
} bmcts_agent_state;

typedef struct bmcts_plan_result {
  uint32_t best_action;    /* GetLastMacroAction() */
  uint32_t num_iterations; /* GetLastNumIterations() */
  uint32_t num_nodes;      /* GetLastNumNodes() */
  uint32_t num_actions;    /* GetNumMotionPrimitives(world) */
  double solution_time_ms; /* GetLastSolutionTime() */
  double return_values[BMCTS_MAX_EGO_ACTIONS];  /* GetLastReturnValues(); NaN = not visited */
  double visit_counts[BMCTS_MAX_EGO_ACTIONS];
  double cost_values[2][BMCTS_MAX_EGO_ACTIONS]; /* GetLastCostValues(): [0] envelope / cost, [1] collision */
  double policy[BMCTS_MAX_EGO_ACTIONS];         /* GetLastPolicySampled().second */
  double expected_risk[2];                      /* GetLastExpectedRisk() */
  double scenario_risk[2];                      /* GetLastScenarioRisk() */
  float action_acc, action_steering;            /* GetLastAction() = Input[acc, steering] */
  int32_t max_tree_depth;
  int32_t root_expanded_actions[BMCTS_MAX_AGENTS];
  int32_t n_traj;
  double trajectory[BMCTS_PLAN_MAX_TRAJ * 5];   /* GetLastTrajectory(): rows [t, x, y, theta, v] */
  uint64_t engine_launches;                     /* kernels launched so far by this planner */
} bmcts_plan_result;

/* mamcts MctsEdgeInfo of the last search tree (behavior_uct_base.hpp:26,32: one record per node,
 * agent and expanded action, down to BehaviorUctBase::MaxExtractionDepth) */
typedef struct bmcts_edge_info {
  uint32_t agent_id;
  uint32_t depth;   /* depth of the node the action leaves from (root = 0) */
  uint32_t action;  /* ego / cooperative agents: macro-action index; agents under hypotheses: the bit
                       pattern of the sampled acceleration (the reference hashes the action value) */
  float weight;     /* share of the node's visits of this agent that took the action */
} bmcts_edge_info;

typedef struct bmcts_planner bmcts_planner;

int bmcts_planner_create(int kind, int device, bmcts_planner** out);
void bmcts_planner_destroy(bmcts_planner* p);
/* BARK ParamServer key ("BehaviorUctBase::Mcts::MaxNumIterations", ...), 1..n values;
 * must be called before the first bmcts_planner_plan */
int bmcts_planner_set_param(bmcts_planner* p, const char* key, const double* values, int n);
/* appends one BehaviorHypothesisIDM to the planner's hypothesis set */
int bmcts_planner_add_hypothesis(bmcts_planner* p, const bmcts_hypothesis* region, int num_samples,
                                 int num_buckets, float buckets_lower, float buckets_upper);
/* the scenario-risk function of BehaviorUCTRiskConstraint, reduced to what the planner calls
 * (behavior_uct_risk_constraint.cpp:51-72): ScenarioRiskFunction::CalculateIntegralValue of
 * every hypothesis region, in hypothesis order (computed with include/bmcts_risk_c.h) */
int bmcts_planner_set_scenario_risk(bmcts_planner* p, const double* integral_per_hypothesis, int n);
/* lane corridors, road polygon and goal of `map` are read; the rest is ignored */
int bmcts_planner_set_map(bmcts_planner* p, const bmcts_scenario* map);
/* root-parallel tree index of this planner (rank): distinct Philox streams per rank */
int bmcts_planner_set_tree_index(bmcts_planner* p, uint32_t tree_index);
/* The behaviour model agent `agent_id` really follows (BARK Agent::GetBehaviorModel), as a
 * BehaviorIDMStochastic parameter region; used when
 * BehaviorUctHypothesis::PredictionSettings::UseTrueBehaviorsAsHypothesis is set
 * (behavior_uct_hypothesis_base.hpp:80-90: the omniscient planner -- every other agent is predicted
 * with its own model, no beliefs).  Kept until overwritten. */
int bmcts_planner_set_true_behavior(bmcts_planner* p, uint32_t agent_id, const bmcts_hypothesis* model);
/* The goal definition of agent `agent_id` (BARK Agent::GetGoalDefinition); NULL forgets it.  The
 * cooperative planner evaluates every agent against its own goal (mcts_state_cooperative.cpp:76-84);
 * an agent without one shares the ego's (the map's).  At most BMCTS_MAX_EXTRA_GOALS distinct
 * definitions besides the ego's. */
int bmcts_planner_set_agent_goal(bmcts_planner* p, uint32_t agent_id, const bmcts_goal* goal);
/* BehaviorUCT*::Plan(delta_time, observed_world) */
int bmcts_planner_plan(bmcts_planner* p, double delta_time, double world_time, uint32_t ego_id,
                       const bmcts_agent_state* agents, int n_agents, bmcts_plan_result* out);
/* BeliefCalculator::BeliefUpdate (belief_calculator/belief_calculator.hpp:23-47) on a
 * hypothesis-based planner handle: one belief update from this observed world without planning
 * (Plan() does the same update itself); bmcts_planner_reset_beliefs = BeliefCalculator::Reset */
int bmcts_planner_update_beliefs(bmcts_planner* p, double world_time, uint32_t ego_id,
                                 const bmcts_agent_state* agents, int n_agents);
int bmcts_planner_reset_beliefs(bmcts_planner* p);
/* GetCurrentBeliefs()[agent_id]; returns the number of hypotheses written (<= max_h) */
int bmcts_planner_get_beliefs(bmcts_planner* p, uint32_t agent_id, double* beliefs, int max_h);
/* flattened root statistics of the last search (per action: visits, sum of returns, sum of
 * cost[0], sum of cost[1]; then iterations, nodes): what root-parallel ranks all-reduce.
 * Returns the number of doubles written. */
int bmcts_planner_root_stats(bmcts_planner* p, double* stats, int max_n);
/* re-decide the action from merged statistics (after the NCCL all-reduce) */
int bmcts_planner_decide_from_root_stats(bmcts_planner* p, const double* stats, int n,
                                         bmcts_plan_result* out);
/* the two calls above with the all-reduce in between: sums the root statistics of the last search
 * over the ranks of `comm` (bmcts_allreduce_root) and re-decides from the merged statistics, so
 * that every rank of a root-parallel Plan() returns the same action */
int bmcts_planner_allreduce_root_stats(bmcts_planner* p, bmcts_comm* comm, bmcts_plan_result* out);
/* GetLastMctsEdgeInfo(): writes up to max_n records, returns how many the last search holds
 * (BehaviorUctBase::ExtractEdgeInfo, default true) */
int bmcts_planner_last_edges(bmcts_planner* p, bmcts_edge_info* edges, int max_n);
/* GetLastMctsStateInfo() (BehaviorUctBase::ExtractStateInfo, default false): per extracted node its
 * depth, alive mask over the search's agent slots and poses [A][4]; returns the number of nodes
 * held, writes up to max_n; *n_agents receives A */
int bmcts_planner_last_states(bmcts_planner* p, uint32_t* depth, uint32_t* agent_ids, float* pose,
                              int max_n, int* n_agents);
/* model of the neural warm start (kind BMCTS_PLANNER_NHEURISTIC_RISK_CONSTRAINT); must be set
 * before the first bmcts_planner_plan.  `user` is passed back to every call. */
int bmcts_planner_set_node_prior(bmcts_planner* p, bmcts_node_prior_fn fn, void* user);
/* length of one observation: (ML::NearestObserver::NNearestAgents + 1) * 4 */
int bmcts_planner_observation_size(bmcts_planner* p);
/* the observer the search applies to its nodes, on an observed world (to build training data
 * and to test models): [x, y, theta, v] of the ego and of its nearest agents, normalised to
 * [0, 1]; needs bmcts_planner_set_map.  Returns the number of floats written or < 0. */
int bmcts_planner_observe(bmcts_planner* p, uint32_t ego_id, const bmcts_agent_state* agents,
                          int n_agents, float* observation, int max_n);
/* calls of the node-prior callback / nodes evaluated so far */
int bmcts_planner_node_prior_stats(bmcts_planner* p, uint64_t* calls, uint64_t* nodes);
/* BehaviorHypothesis::ParameterSamplingProbability of hypothesis h */
double bmcts_planner_parameter_sampling_probability(bmcts_planner* p, int h);
const char* bmcts_planner_last_error(const bmcts_planner* p);

#ifdef __cplusplus
}
#endif
#endif /* BMCTS_PLANNER_C_H_ */
