// planner_handle.hpp -- the object behind the opaque bmcts_planner handle.
#pragma once

#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "behavior_uct.hpp"
#include "bmcts_planner_c.h"

struct bmcts_planner {
  int kind = 0;
  int device = 0;
  unsigned tree_index = 0;
  std::shared_ptr<bark_mcts_b200::Params> params = std::make_shared<bark_mcts_b200::Params>();
  std::vector<bark_mcts_b200::BehaviorHypothesisIDM> hypotheses;
  std::map<uint32_t, bmcts_goal> agent_goals;  // agent id -> its own goal definition
  std::map<uint32_t, bmcts_hypothesis> true_behaviors;  // agent id -> its own behaviour model
  std::vector<double> risk_integrals;  // scenario-risk integral per hypothesis (may be empty)
  std::shared_ptr<bark_mcts_b200::MapGeometry> map;
  // how the leaf evaluator is made; the product default is the CUDA engine
  std::function<std::shared_ptr<bark_mcts_b200::LeafEvaluator>()> evaluator_factory;
  std::shared_ptr<bark_mcts_b200::CallbackPriorProvider> node_prior;  // neural warm start
  std::shared_ptr<bark_mcts_b200::BehaviorUCTBase> planner;  // built lazily at the first plan
  std::string err;
};
