// test_comm.cpp -- the collective of the path through the C ABI alone (no Python, no torch):
// bmcts_comm_* + bmcts_allreduce_root (include/bmcts_c.h), and a root-parallel
// Plan() of two BehaviorUCTHypothesis planners merged with bmcts_planner_allreduce_root_stats.
// One rank per visible GPU, at most two (ranks are threads of this process; on a one-GPU box the
// communicator has a single rank and the all-reduce is the identity).
// Built and run by tests/test_root_parallel.py against libbmcts.so; prints "ok" on success.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include <cuda_runtime_api.h>

#include "bmcts_planner_c.h"

#define CHECK(c)                                                        \
  do {                                                                  \
    if (!(c)) {                                                         \
      std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); \
      std::exit(1);                                                     \
    }                                                                   \
  } while (0)

static void Map(bmcts_scenario* m) {  // MakeXodrMapOneRoadTwoLanes
  std::memset(m, 0, sizeof(*m));
  m->n_corridors = 2;
  const float ys[2] = {-1.75f, -4.75f};
  for (int c = 0; c < 2; ++c) {
    bmcts_corridor& k = m->corridors[c];
    k.n_pts = 2;
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
  const float gx[4] = {146, 154, 154, 146}, gy[4] = {2.25f, 2.25f, -5.75f, -5.75f};
  for (int i = 0; i < 4; ++i) {
    m->goal.px[i] = gx[i];
    m->goal.py[i] = gy[i];
  }
}

struct RankOut {
  std::vector<double> local, merged;
  bmcts_plan_result result;
};

static void Rank(int rank, int nranks, int device, const void* id, RankOut* out) {
  bmcts_comm* comm = nullptr;
  CHECK(bmcts_comm_create(device, nranks, rank, id, &comm) == BMCTS_OK);
  CHECK(bmcts_comm_nranks(comm) == nranks);
  // plain vector first
  double v[3] = {1.0 + rank, 10.0 * (rank + 1), 0.5};
  CHECK(bmcts_allreduce_root(comm, v, 3) == BMCTS_OK);
  double want[3] = {0, 0, 0};
  for (int r = 0; r < nranks; ++r) {
    want[0] += 1.0 + r;
    want[1] += 10.0 * (r + 1);
    want[2] += 0.5;
  }
  CHECK(v[0] == want[0] && v[1] == want[1] && v[2] == want[2]);
  // a planner per rank, distinct tree indices, merged decision
  bmcts_planner* p = nullptr;
  CHECK(bmcts_planner_create(BMCTS_PLANNER_HYPOTHESIS, device, &p) == BMCTS_OK);
  auto set = [&](const char* key, std::vector<double> val) {
    CHECK(bmcts_planner_set_param(p, key, val.data(), (int)val.size()) == BMCTS_OK);
  };
  set("BehaviorUctBase::Mcts::MaxNumIterations", {600});
  set("BehaviorUctBase::Mcts::MaxSearchTime", {1e9});
  set("BehaviorUctBase::Mcts::RandomHeuristic::MaxNumIterations", {20});
  set("BehaviorUctBase::EgoBehavior::AccelerationInputs", {0, 1, 4, -1, -8});
  set("BehaviorUctBase::EgoBehavior::AddLaneChangeActions", {0});
  set("BehaviorUctBase::Mcts::UctStatistic::ProgressiveWidening::Alpha", {0.5});
  set("BehaviorUctBase::Mcts::UctStatistic::ProgressiveWidening::K", {0.4});
  set("BehaviorUctBase::Mcts::State::GoalReward", {200});
  set("BehaviorUctBase::Mcts::State::CollisionReward", {-2000});
  for (int h = 0; h < 2; ++h) {
    bmcts_hypothesis hy;
    std::memset(&hy, 0, sizeof(hy));
    const float fixed[6] = {0, 0, 1, 15, 1, 0};
    for (int m = 0; m < 6; ++m) hy.lo[m] = hy.hi[m] = fixed[m];
    hy.lo[0] = h == 0 ? 1.0f : 1.5f;
    hy.hi[0] = h == 0 ? 1.5f : 3.0f;
    hy.acc_lower_bound = -5;
    hy.acc_upper_bound = 8;
    hy.max_velocity = 50;
    hy.exponent = 4;
    CHECK(bmcts_planner_add_hypothesis(p, &hy, 1000, 100, -10.0f, 10.0f) == BMCTS_OK);
  }
  bmcts_scenario map;
  Map(&map);
  CHECK(bmcts_planner_set_map(p, &map) == BMCTS_OK);
  CHECK(bmcts_planner_set_tree_index(p, (uint32_t)rank) == BMCTS_OK);
  bmcts_agent_state agents[2] = {{1, 3.0f, -1.75f, 0.0f, 5.0f, 4.0f, 2.0f, 1.25f, 0.0f},
                                 {2, 9.0f, -1.75f, 0.0f, 1.0f, 4.0f, 2.0f, 1.25f, 0.0f}};
  bmcts_plan_result res;
  CHECK(bmcts_planner_plan(p, 0.2, 0.0, 1, agents, 2, &res) == BMCTS_OK);
  CHECK(res.num_iterations == 600 && res.engine_launches > 0);
  out->local.resize(64);
  const int n = bmcts_planner_root_stats(p, out->local.data(), 64);
  CHECK(n > 0);
  out->local.resize(n);
  CHECK(bmcts_planner_allreduce_root_stats(p, comm, &out->result) == BMCTS_OK);
  out->merged.resize(64);
  out->merged.resize(bmcts_planner_root_stats(p, out->merged.data(), 64));
  bmcts_planner_destroy(p);
  bmcts_comm_destroy(comm);
}

int main() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    std::fprintf(stderr, "no CUDA device\n");
    return 2;
  }
  const int nranks = ndev >= 2 ? 2 : 1;
  unsigned char id[BMCTS_COMM_ID_BYTES];
  CHECK(bmcts_comm_unique_id(id) == BMCTS_OK);
  std::vector<RankOut> outs(nranks);
  std::vector<std::thread> th;
  for (int r = 0; r < nranks; ++r) th.emplace_back(Rank, r, nranks, r, (const void*)id, &outs[r]);
  for (auto& t : th) t.join();
  // merged statistics = sum of the ranks' statistics, the same on every rank; same decision
  for (int r = 0; r < nranks; ++r) {
    CHECK(outs[r].merged.size() == outs[0].local.size());
    for (size_t i = 0; i < outs[r].merged.size(); ++i) {
      double sum = 0.0;
      for (int q = 0; q < nranks; ++q) sum += outs[q].local[i];
      CHECK(std::fabs(outs[r].merged[i] - sum) <= 1e-9 * std::fabs(sum));
    }
    CHECK(outs[r].result.best_action == outs[0].result.best_action);
    CHECK(outs[r].result.num_iterations == 600u * (unsigned)nranks);
    CHECK(outs[r].result.action_acc < 0.0f);   // a slow car 2 m ahead: brake
    CHECK(outs[r].result.n_traj >= 2);
  }
  if (nranks == 2) CHECK(outs[0].local != outs[1].local);  // distinct Philox streams per rank
  std::printf("ranks %d ok\n", nranks);
  return 0;
}
