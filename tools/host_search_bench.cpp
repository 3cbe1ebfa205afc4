// host_search_bench.cpp -- cost of the host tree search alone (select / fill / backup per
// iteration) with a leaf evaluator that answers instantly.  Development tool: shows what the
// host side of Plan() can sustain when the GPU is not the limit.
//   g++ -O3 -std=c++17 -Iinclude -Iplanner-mcts_b200/host tools/host_search_bench.cpp \
//       planner-mcts_b200/host/behavior_uct.cpp planner-mcts_b200/csrc/scenario.cpp -lpthread
//   ./a.out <agents> <hypotheses> <iterations> <wave> <risk 0|1> [trees]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>

#include "behavior_uct.hpp"

using namespace bark_mcts_b200;

namespace {

uint64_t Mix(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  return x ^ (x >> 33);
}

// next state = constant-velocity advance; planned accelerations take few distinct values so
// that joint actions repeat about as often as with the IDM hypotheses
class InstantEvaluator : public LeafEvaluator {
 public:
  int Configure(const bmcts_config&, const bmcts_scenario& scn) override {
    n_prob_ = (size_t)scn.n_agents * scn.n_hypotheses;
    return BMCTS_OK;
  }
  int ExpandRollout(WaveBuffers& w, uint64_t, uint32_t, uint64_t first) override {
    for (int k = 0; k < w.n; ++k) {
      for (int a = 0; a < w.A; ++a) {
        const size_t e = w.at(a, k);
        for (int f = 0; f < 4; ++f) w.next_pose[e * 4 + f] = w.pose[e * 4 + f];
        w.next_pose[e * 4] += 0.5f * w.pose[e * 4 + 3];
        w.reward[e] = 0.0f;
        w.last_action[e] = 0.25f * (float)(Mix(w.draw_id[e] * 31 + w.hyp_id[e]) % 24) - 3.0f;
      }
      w.next_alive[k] = w.alive[k];
      w.flags[k] = 0;
      w.cost[k] = 0.0f;
      w.cost[(size_t)w.n + k] = 0.0f;
      const uint64_t r = Mix(first + k);
      w.result[k] = bmcts_rollout_result{};
      w.result[k].ret_ego = -50.0f + (float)(r % 100);
      w.result[k].cost0 = (float)((r >> 8) % 100) * 0.01f;
      w.result[k].n_steps = 20;
    }
    calls_++;
    return BMCTS_OK;
  }
  int Likelihood(const std::vector<float>&, uint32_t, const std::vector<float>&, int, int, float,
                 float, uint64_t, std::vector<float>* prob) override {
    prob->assign(n_prob_, 0.5f);
    return BMCTS_OK;
  }
  uint64_t LaunchCount() const override { return calls_; }
  std::string LastError() const override { return ""; }

 private:
  uint64_t calls_ = 0;
  size_t n_prob_ = 0;
};

std::shared_ptr<MapGeometry> Road(int lanes, float length) {
  auto m = std::make_shared<MapGeometry>();
  m->n_corridors = lanes;
  for (int c = 0; c < lanes; ++c) {
    bmcts_corridor& k = m->corridors[c];
    k.n_pts = 2;
    k.px[0] = 0.0f;
    k.px[1] = length;
    k.py[0] = k.py[1] = -1.75f - 3.5f * c;
    k.half_width = 1.75f;
    k.left = c - 1;
    k.right = c + 1 < lanes ? c + 1 : -1;
  }
  m->n_road_pts = 4;
  const float rx[4] = {0, length, length, 0}, ry[4] = {0, 0, -3.5f * lanes, -3.5f * lanes};
  for (int i = 0; i < 4; ++i) {
    m->road_x[i] = rx[i];
    m->road_y[i] = ry[i];
  }
  m->goal.kind = BMCTS_GOAL_POLYGON;
  m->goal.n_pts = 4;
  const float gx[4] = {length - 20, length - 10, length - 10, length - 20};
  const float gy[4] = {0, 0, -3.5f, -3.5f};
  for (int i = 0; i < 4; ++i) {
    m->goal.px[i] = gx[i];
    m->goal.py[i] = gy[i];
  }
  return m;
}

}  // namespace

namespace bark_mcts_b200 {
// this tool never touches a GPU: the product factory resolves to the instant evaluator
std::unique_ptr<LeafEvaluator> MakeDefaultEvaluator(int) {
  return std::unique_ptr<LeafEvaluator>(new InstantEvaluator());
}
}  // namespace bark_mcts_b200

int main(int argc, char** argv) {
  const int A = argc > 1 ? std::atoi(argv[1]) : 10;
  const int H = argc > 2 ? std::atoi(argv[2]) : 8;
  const long iterations = argc > 3 ? std::atol(argv[3]) : 32768;
  const long wave = argc > 4 ? std::atol(argv[4]) : 1024;
  const bool risk = argc > 5 && std::atoi(argv[5]) != 0;
  const long trees = argc > 6 ? std::atol(argv[6]) : 1;

  auto params = std::make_shared<Params>();
  params->SetInt("BehaviorUctBase::Mcts::MaxNumIterations", iterations);
  params->SetInt("BehaviorUctBase::Mcts::MaxNumNodes", iterations + 10);
  params->SetInt("BehaviorUctBase::Mcts::MaxSearchTime", 100000000);
  params->SetInt("BehaviorUctBase::Mcts::Engine::WaveSize", wave);
  params->SetInt("BehaviorUctBase::MaxNearestAgents", A - 1);
  params->SetInt("BehaviorUctBase::Mcts::NumParallelMcts", trees);
  params->SetBool("BehaviorUctBase::Mcts::UseMultiThreading", trees > 1);
  params->SetListFloat("BehaviorUctBase::EgoBehavior::AccelerationInputs", {0, 1, 4, -1, -8});
  if (risk) {
    params->SetBool("BehaviorUctBase::Mcts::State::SplitSafeDistCollision", true);
    params->SetListFloat("BehaviorUctRiskConstraint::DefaultAvailableRisk", {0.1, 0.0});
    params->SetBool("BehaviorUctBase::Mcts::HypothesisStatistic::CostBasedActionSelection", true);
  }

  ObservedWorld world;
  world.map = Road(3, 600.0f);
  world.ego_id = 1;
  for (int a = 0; a < A; ++a) {
    Agent ag;
    ag.id = 1 + a;
    ag.state = {0.0, 30.0 + 12.0 * (a / 3), -1.75 - 3.5 * (a % 3), 0.0, 8.0 + (a % 5)};
    world.agents.push_back(ag);
  }
  std::vector<BehaviorHypothesisIDM> hyps;
  for (int h = 0; h < H; ++h) {
    BehaviorHypothesisIDM hy;
    const float fixed[6] = {1.5f, 2, 1.7f, 15, 1.67f, 0};
    for (int m = 0; m < 6; ++m) hy.region.lo[m] = hy.region.hi[m] = fixed[m];
    hy.region.lo[0] = 0.5f + 3.0f * h / H;
    hy.region.hi[0] = 0.5f + 3.0f * (h + 1) / H;
    hy.region.acc_lower_bound = -5;
    hy.region.acc_upper_bound = 8;
    hy.region.max_velocity = 50;
    hy.region.exponent = 4;
    hy.num_samples = 100;
    hyps.push_back(hy);
  }
  auto make = []() { return std::shared_ptr<LeafEvaluator>(new InstantEvaluator()); };
  std::unique_ptr<BehaviorUCTBase> planner;
  if (risk)
    planner.reset(new BehaviorUCTRiskConstraint(params, hyps, nullptr, make()));
  else
    planner.reset(new BehaviorUCTHypothesis(params, hyps, make()));
  planner->set_evaluator_factory(make);
  planner->Plan(0.2, world);  // warm-up
  const auto t0 = std::chrono::steady_clock::now();
  planner->Plan(0.2, world);
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  const double it = (double)planner->GetLastNumIterations();
  std::printf("A=%d H=%d wave=%ld trees=%ld %s: %.0f iterations in %.1f ms = %.2f us / iteration "
              "(%.3g iterations/s), %u nodes, depth %d\n",
              A, H, wave, trees, risk ? "risk" : "hypothesis", it, ms, 1000.0 * ms / it * trees,
              it / ms * 1000.0, planner->GetLastNumNodes(), planner->GetLastMaxTreeDepth());
  return 0;
}
