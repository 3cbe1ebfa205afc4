// test_statistics.cpp -- unit tests of the node statistics of the host tree search
// (planner-mcts_b200/host/statistics.hpp), header only: no engine, no GPU.
//
//  * CostConstrainedStatistic::GreedyPolicy with TWO cost constraints
//    (/root/reference/bark_mcts/models/behavior/behavior_uct_risk_constraint.hpp:145-181 calls
//    greedy_policy on a statistic that carries a vector of constraints, defaults {0.1, 0}):
//    the policy must meet both, mixing up to three actions, and minimise the lambda-weighted
//    violation when they cannot be met;
//  * UctStatistic progressive widening.
// Built by oracle/Makefile, run by tests/test_cpp_api.py; prints "ok" and exits 0 on success.
#include <cmath>
#include <cstdio>
#include <random>

#include "statistics.hpp"

using namespace bark_mcts_b200;

#define CHECK(c)                                                        \
  do {                                                                  \
    if (!(c)) {                                                         \
      std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); \
      return 1;                                                         \
    }                                                                   \
  } while (0)

// a statistic whose actions have the given mean return / envelope cost / collision cost
static CostConstrainedStatistic Make(int n, const double* ret, const double* c0, const double* c1,
                                     const std::vector<double>& constraints, double visits = 100) {
  MctsParameters m;
  m.cost_constrained_statistic.REWARD_LOWER_BOUND = 0.0;
  m.cost_constrained_statistic.REWARD_UPPER_BOUND = 1.0;
  m.cost_constrained_statistic.COST_CONSTRAINTS = constraints;
  CostConstrainedStatistic s;
  s.Init(n, 2, m);
  for (int a = 0; a < n; ++a) {
    s.reward_statistic().MergeRaw(a, visits, visits * ret[a]);
    s.cost_statistic(0).MergeRaw(a, visits, visits * c0[a]);
    s.cost_statistic(1).MergeRaw(a, visits, visits * c1[a]);
  }
  s.set_constraints(constraints);
  return s;
}

static double Prob(const Policy& p, unsigned a) {
  auto it = p.find(a);
  return it == p.end() ? 0.0 : it->second;
}

int main() {
  std::mt19937 gen(1);
  const std::vector<double> lambdas = {0.0, 0.0};  // plain reward ordering, large filter: all near
  const double filter = 1e9;
  {
    // Channel 1 binds.  Action 0 is the best but violates both constraints; action 1 is free of
    // envelope risk but carries collision hazard; action 2 is safe on both and poor.  The
    // single-constraint mix (first channel only) puts 2/3 on action 0 and 1/3 on action 1:
    // envelope cost 0.1 as required, collision cost 2/3 * 0.3 + 1/3 * 0.15 = 0.25 > 0.05.
    const double ret[3] = {1.0, 0.6, 0.1}, c0[3] = {0.15, 0.0, 0.0}, c1[3] = {0.3, 0.15, 0.0};
    CostConstrainedStatistic s = Make(3, ret, c0, c1, {0.1, 0.05});
    auto sampled = s.GreedyPolicy(0.0, filter, lambdas, gen);
    const Policy& pol = sampled.second;
    std::vector<double> e = s.ExpectedPolicyCost(pol);
    CHECK(e[0] <= 0.1 + 1e-12 && e[1] <= 0.05 + 1e-12);  // both constraints met
    const double old_mix_c1 = 2.0 / 3.0 * 0.3 + 1.0 / 3.0 * 0.15;
    CHECK(old_mix_c1 > 0.05);                             // ... which the two-action mix does not
    double sum = 0.0, value = 0.0;
    for (const auto& kv : pol) {
      sum += kv.second;
      value += kv.second * ret[kv.first];
    }
    CHECK(std::fabs(sum - 1.0) < 1e-12);
    // the LP optimum: maximise return subject to both constraints.  Vertices: channel 1 tight
    // on {0,2}: p0 = 1/6, return 0.25; on {1,2}: p1 = 1/3, return 0.267; both tight on {0,1,2}
    // needs p1 < 0.  Best: one third of action 1, two thirds of action 2.
    CHECK(std::fabs(Prob(pol, 1) - 1.0 / 3.0) < 1e-9 && std::fabs(Prob(pol, 2) - 2.0 / 3.0) < 1e-9);
    CHECK(std::fabs(value - (0.6 / 3.0 + 0.2 / 3.0)) < 1e-9);
    CHECK(pol.count((unsigned)sampled.first) == 1);
  }
  {
    // both constraints tight at once: three actions in the mix
    const double ret[3] = {1.0, 0.9, 0.0}, c0[3] = {0.3, 0.0, 0.0}, c1[3] = {0.0, 0.3, 0.0};
    CostConstrainedStatistic s = Make(3, ret, c0, c1, {0.1, 0.1});
    const Policy pol = s.GreedyPolicy(0.0, filter, lambdas, gen).second;
    std::vector<double> e = s.ExpectedPolicyCost(pol);
    CHECK(pol.size() == 3);
    CHECK(std::fabs(e[0] - 0.1) < 1e-12 && std::fabs(e[1] - 0.1) < 1e-12);
    CHECK(std::fabs(Prob(pol, 0) - 1.0 / 3.0) < 1e-9 && std::fabs(Prob(pol, 1) - 1.0 / 3.0) < 1e-9);
  }
  {
    // the maximising action is feasible: no mixing, no random number consumed
    const double ret[2] = {1.0, 0.5}, c0[2] = {0.05, 0.0}, c1[2] = {0.0, 0.0};
    CostConstrainedStatistic s = Make(2, ret, c0, c1, {0.1, 0.0});
    std::mt19937 g1(7), g2(7);
    auto sampled = s.GreedyPolicy(0.0, filter, lambdas, g1);
    CHECK(sampled.first == 0 && sampled.second.size() == 1 && g1() == g2());
  }
  {
    // infeasible whatever the mix (every action has collision cost above the limit 0): the
    // policy minimises the lambda-weighted violation -- here the least hazardous action
    const double ret[3] = {1.0, 0.8, 0.2}, c0[3] = {0.0, 0.0, 0.0}, c1[3] = {0.5, 0.2, 0.3};
    CostConstrainedStatistic s = Make(3, ret, c0, c1, {0.1, 0.0});
    auto sampled = s.GreedyPolicy(0.0, filter, {1.0, 1.0}, gen);
    CHECK(sampled.first == 1 && sampled.second.size() == 1);
  }
  {
    // one cost channel (non-split costs): the classical two-action mix at the constraint
    MctsParameters m;
    m.cost_constrained_statistic.COST_CONSTRAINTS = {0.1};
    CostConstrainedStatistic s;
    s.Init(2, 1, m);
    s.reward_statistic().MergeRaw(0, 50, 50 * 90.0);
    s.reward_statistic().MergeRaw(1, 50, 50 * 10.0);
    s.cost_statistic(0).MergeRaw(0, 50, 50 * 0.3);
    s.cost_statistic(0).MergeRaw(1, 50, 50 * 0.0);
    s.set_constraints({0.1});
    const Policy pol = s.GreedyPolicy(0.0, filter, {0.0}, gen).second;
    CHECK(std::fabs(Prob(pol, 0) - 1.0 / 3.0) < 1e-9 && std::fabs(Prob(pol, 1) - 2.0 / 3.0) < 1e-9);
  }
  {
    // UctStatistic progressive widening: #expanded <= K N^alpha (reference defaults 1, 0.1)
    UctStatistic u;
    u.Init(5, -1000, 100, 0.7, 0.9, true, 1.0, 0.1);
    int expanded_at[4] = {0, 0, 0, 0};
    for (int visit = 0; visit < 1100; ++visit) {
      const int a = u.ChooseNextAction(gen);
      u.ReleasePending(a);
      u.UpdateStatistic(a, 0.0, (double)a);
      int n = 0;
      for (const auto& pr : u.pairs()) n += pr.expanded;
      if (visit == 0) expanded_at[0] = n;
      if (visit == 1) expanded_at[1] = n;
      if (visit == 1000) expanded_at[2] = n;
      if (visit == 1099) expanded_at[3] = n;
    }
    CHECK(expanded_at[0] == 1 && expanded_at[1] == 2 && expanded_at[2] == 2 && expanded_at[3] == 3);
  }
  std::printf("ok\n");
  return 0;
}
