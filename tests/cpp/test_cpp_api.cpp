// test_cpp_api.cpp -- the C++ planner surface used the way the reference's C++ tests use theirs
// (/root/reference/bark_mcts/models/behavior/tests/behavior_uct_hypothesis_test.cc:215-240,
//  behavior_uct_cooperative_test.cc:147-165): construct from parameters + hypotheses,
// Plan(delta_time, observed_world), read the debug infos, Clone().
// Built against liboracle_planner.so (CPU leaf evaluator) by oracle/Makefile and run by
// tests/test_cpp_api.py; prints "ok" and exits 0 on success.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>

#include "behavior_uct.hpp"

using namespace bark_mcts_b200;

#define CHECK(c)                                                        \
  do {                                                                  \
    if (!(c)) {                                                         \
      std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); \
      return 1;                                                         \
    }                                                                   \
  } while (0)

static std::shared_ptr<MapGeometry> OneRoadTwoLanes() {  // MakeXodrMapOneRoadTwoLanes
  auto m = std::make_shared<MapGeometry>();
  m->n_corridors = 2;
  const float ys[2] = {-1.75f, -4.75f};
  for (int c = 0; c < 2; ++c) {
    bmcts_corridor& k = m->corridors[c];
    k.n_pts = 2;
    k.px[0] = 0.0f;
    k.px[1] = 300.0f;
    k.py[0] = k.py[1] = ys[c];
    k.half_width = c == 0 ? 1.75f : 1.25f;
    k.left = c == 0 ? -1 : 0;
    k.right = c == 0 ? 1 : -1;
  }
  m->n_road_pts = 4;
  const float rx[4] = {0, 300, 300, 0}, ry[4] = {0, 0, -6, -6};
  for (int i = 0; i < 4; ++i) {
    m->road_x[i] = rx[i];
    m->road_y[i] = ry[i];
  }
  m->goal.kind = BMCTS_GOAL_POLYGON;
  m->goal.n_pts = 4;
  const float gx[4] = {56, 64, 64, 56}, gy[4] = {2.25f, 2.25f, -5.75f, -5.75f};
  for (int i = 0; i < 4; ++i) {
    m->goal.px[i] = gx[i];
    m->goal.py[i] = gy[i];
  }
  return m;
}

static BehaviorHypothesisIDM Hypothesis(float lo, float hi) {  // make_params_hypothesis
  BehaviorHypothesisIDM h;
  const float fixed[6] = {0, 0, 1, 15, 1, 0};  // headway, spacing, max acc, v0, comfort braking, coolness
  for (int m = 0; m < 6; ++m) h.region.lo[m] = h.region.hi[m] = fixed[m];
  h.region.lo[0] = lo;
  h.region.hi[0] = hi;
  h.region.acc_lower_bound = -5;
  h.region.acc_upper_bound = 8;
  h.region.min_velocity = 0;
  h.region.max_velocity = 50;
  h.region.exponent = 4;
  h.num_samples = 500;
  return h;
}

int main() {
  auto params = std::make_shared<Params>();
  params->SetInt("BehaviorUctBase::Mcts::MaxNumIterations", 400);
  params->SetInt("BehaviorUctBase::Mcts::MaxSearchTime", 1000000);
  params->SetListFloat("BehaviorUctBase::EgoBehavior::AccelerationInputs", {0, 1, 4, -1, -8});
  params->SetBool("BehaviorUctBase::EgoBehavior::AddLaneChangeActions", false);
  // behavior_uct_hypothesis_test.cc:253-254 (the defaults open a third ego action after 1024 visits)
  params->SetReal("BehaviorUctBase::Mcts::UctStatistic::ProgressiveWidening::Alpha", 0.5);
  params->SetReal("BehaviorUctBase::Mcts::UctStatistic::ProgressiveWidening::K", 0.4);
  params->SetReal("BehaviorUctBase::Mcts::State::GoalReward", 200.0);
  params->SetReal("BehaviorUctBase::Mcts::State::CollisionReward", -2000.0);

  ObservedWorld world;
  world.map = OneRoadTwoLanes();
  world.ego_id = 1;
  Agent ego;
  ego.id = 1;
  ego.state = {0.0, 3.0, -1.75, 0.0, 5.0};
  world.agents.push_back(ego);

  std::vector<BehaviorHypothesisIDM> hyps = {Hypothesis(1.0f, 1.5f), Hypothesis(1.5f, 3.0f)};
  std::shared_ptr<LeafEvaluator> evaluator(MakeDefaultEvaluator(0));  // the CPU oracle in this build

  // free road: accelerate (behavior_uct_hypothesis_test.cc:239)
  BehaviorUCTHypothesis planner(params, hyps, evaluator);
  Trajectory traj = planner.Plan(0.2, world);
  CHECK(planner.GetLastAction().acc >= 0.0);
  CHECK(planner.GetLastNumIterations() == 400);
  CHECK(traj.size() >= 2 && std::fabs(traj.back().t - 0.2) < 1e-6 && traj.back().x > traj.front().x);
  CHECK(!planner.GetLastReturnValues().empty());
  CHECK(planner.GetLastMacroAction() < 5);
  CHECK(std::fabs(hyps[0].ParameterSamplingProbability() - 2.0) < 1e-4);  // 1 / (1.5 - 1.0)

  // Clone(): an independent planner with the same parameters and hypotheses
  std::shared_ptr<BehaviorUCTBase> clone = planner.Clone();
  Trajectory traj2 = clone->Plan(0.2, world);
  CHECK(traj2.size() == traj.size());
  CHECK(clone->GetLastMacroAction() == planner.GetLastMacroAction());

  // a slow agent 2 m ahead: brake (behavior_uct_hypothesis_test.cc:279)
  Agent other;
  other.id = 2;
  other.state = {0.0, 3.0 + 2.0 + 4.0, -1.75, 0.0, 1.0};
  world.agents.push_back(other);
  params->SetBool("BehaviorUctBase::Mcts::HypothesisStatistic::CostBasedActionSelection", true);
  params->SetInt("BehaviorUctBase::Mcts::MaxNumIterations", 1500);
  BehaviorUCTHypothesis braking(params, hyps, evaluator);
  braking.Plan(0.2, world);
  CHECK(braking.GetLastAction().acc < 0.0);
  Beliefs beliefs = braking.GetCurrentBeliefs();
  CHECK(beliefs.size() == 1 && beliefs.begin()->second.size() == 2);  // belief_calculator_test.cc:91-114

  // cooperative planner on the same world (behavior_uct_cooperative_test.cc:187)
  params->SetReal("BehaviorUctBase::Mcts::State::CooperationFactor", 0.75);
  BehaviorUCTCooperative coop(params, evaluator);
  coop.Plan(0.2, world);
  CHECK(coop.GetLastAction().acc < 0.0);

  // BeliefCalculator (belief_calculator_test.cc:97-123): beliefs without a planner
  {
    BeliefCalculator calc(params, hyps, evaluator);
    calc.BeliefUpdate(world);
    Beliefs b = calc.GetBeliefs();
    CHECK(b.size() == 1 && b.begin()->first == 2u && b.begin()->second.size() == 2);
    CHECK(std::fabs(b.begin()->second[0] + b.begin()->second[1] - 1.0) < 1e-12);
    CHECK(calc.GetBehaviorHypotheses().size() == 2);
    calc.Reset();
    CHECK(calc.GetBeliefs().empty());
  }

  // root-belief sampling of a tree (alias tables over the beliefs): frequencies follow the
  // beliefs, a hypothesis with belief zero is never drawn, a fixed hypothesis is kept
  {
    MctsParameters mp;
    HypothesisBeliefTracker tracker(mp);
    tracker.belief_update({{7u, {0.1, 0.6, 0.3, 0.0}}, {9u, {0.25, 0.25, 0.25, 0.25}}});
    const std::vector<double> want = tracker.get_beliefs().at(7u);
    CHECK(std::fabs(want[0] + want[1] + want[2] + want[3] - 1.0) < 1e-12 && want[3] == 0.0);
    tracker.update_fixed_hypothesis_set({{9u, 2u}});
    auto sampler = tracker.SamplerForSlots({1u, 7u, 9u, 11u}, 4, 3);  // slot 3: no belief yet
    std::vector<uint8_t> hyp(4, 0);
    std::vector<long> count7(4, 0), count11(4, 0);
    const long draws = 400000;
    for (long i = 0; i < draws; ++i) {
      sampler(&hyp);
      count7[hyp[1]]++;
      CHECK(hyp[2] == 2);
      count11[hyp[3]]++;
    }
    CHECK(count7[3] == 0);
    for (int h = 0; h < 4; ++h) {
      CHECK(std::fabs((double)count7[h] / draws - want[h]) < 0.005);
      CHECK(std::fabs((double)count11[h] / draws - 0.25) < 0.005);  // uniform without a belief
    }
  }

  std::printf("ok\n");
  return 0;
}
