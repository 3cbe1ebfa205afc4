// world.hpp -- POD restatement of the BARK types the BehaviorUCT* planners touch.
//
// BARK (github.com/bark-simulator/bark @0c8f1b5) is not available, so the
// planner surface is kept (class names, Plan(delta_time, observed_world), the
// ParamServer keys) over plain structs: an ObservedWorld is the agents' states
// plus the map reduced to what the path reads (lane corridors, road-corridor
// polygon, ego goal).  Paths cited are relative to
// /root/reference/bark_mcts/models/behavior/.
#pragma once

#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "bmcts_c.h"

namespace bark_mcts_b200 {

using AgentId = uint32_t;

// BARK dynamic::State [t, x, y, theta, v] (tests/test_helpers.cpp:71-72)
struct State {
  double t = 0, x = 0, y = 0, theta = 0, v = 0;
};
using Trajectory = std::vector<State>;

// BARK Action variant restricted to what the path uses: the longitudinal acceleration
// (Continuous1DAction of IDM agents) or Input[acc, steering] of the macro actions.
struct Action {
  double acc = 0.0;
  double steering = 0.0;
};

struct Agent {
  AgentId id = 0;
  State state;
  float length = 4.0f, width = 2.0f, rear = 1.25f;  // CarRectangle
  Action last_action;  // GetStateInputHistory().back().second
  // the agent's own behaviour model (BARK Agent::GetBehaviorModel) as a BehaviorIDMStochastic
  // parameter region, if the simulation knows it: what the omniscient planners use as this agent's
  // one hypothesis (BehaviorUctHypothesis::PredictionSettings::UseTrueBehaviorsAsHypothesis,
  // behavior_uct_hypothesis_base.hpp:80-90)
  bool has_behavior_model = false;
  bmcts_hypothesis behavior_model{};
  // the agent's own goal definition (BARK Agent::GetGoalDefinition), if it has one: the
  // cooperative planner evaluates every agent against its own goal
  // (mcts_state_cooperative.cpp:76-84); without one the agent shares the ego's
  bool has_goal = false;
  bmcts_goal goal{};
};

// Map reduced to the path's inputs: only corridors, road polygon and goal of this
// bmcts_scenario are read; agents, hypotheses and actions are filled by the planner.
using MapGeometry = bmcts_scenario;

struct ObservedWorld {
  double world_time = 0.0;
  AgentId ego_id = 0;
  std::vector<Agent> agents;  // ego included
  std::shared_ptr<const MapGeometry> map;

  const Agent* GetAgent(AgentId id) const {
    for (const auto& a : agents)
      if (a.id == id) return &a;
    return nullptr;
  }
  const Agent* GetEgoAgent() const { return GetAgent(ego_id); }
};

// BehaviorHypothesisIDM (hypothesis/idm/hypothesis_idm.hpp:26-50): parameter region of
// BehaviorIDMStochastic + the likelihood histogram settings (hypothesis_idm.cpp:27-38)
struct BehaviorHypothesisIDM {
  bmcts_hypothesis region{};
  int num_samples = 10000;
  int num_buckets = 100;
  float buckets_lower_bound = -10.0f;
  float buckets_upper_bound = 10.0f;

  // BehaviorHypothesis::ParameterSamplingProbability (hypothesis/behavior_hypothesis.hpp:34):
  // 1 / volume of the parameter region (degenerate FixedValue dimensions do not count)
  double ParameterSamplingProbability() const {
    double area = 1.0;
    for (int m = 0; m < BMCTS_NUM_IDM_PARAMS; ++m) {
      double w = (double)region.hi[m] - (double)region.lo[m];
      if (w > 0.0) area *= w;
    }
    return 1.0 / area;
  }
};

using Policy = std::unordered_map<unsigned, double>;  // mcts::Policy
using Beliefs = std::map<AgentId, std::vector<double>>;

}  // namespace bark_mcts_b200
