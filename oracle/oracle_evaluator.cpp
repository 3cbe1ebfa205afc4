// oracle_evaluator.cpp -- TEST INFRASTRUCTURE: a LeafEvaluator backed by the CPU oracle, and
// oracle_planner_create(), which builds the SAME host tree search and BehaviorUCT* planners
// (planner-mcts_b200/host/*) on top of it.  Used by tests/ (search parity: GPU-backed vs
// CPU-backed planner at equal iteration counts) and by bench.py's CPU arm (the reference's
// planner structure -- sequential MCTS with CPU rollouts -- timed on the host cores).
// Nothing in the product links this file.
#include <cstring>
#include <memory>

#include "bmcts_oracle.h"
#include "evaluator.hpp"
#include "planner_handle.hpp"

namespace bark_mcts_b200 {

namespace {

class OracleLeafEvaluator : public LeafEvaluator {
 public:
  int Configure(const bmcts_config& cfg, const bmcts_scenario& scn) override {
    cfg_ = cfg;
    scn_ = scn;
    return BMCTS_OK;
  }
  int ExpandRollout(WaveBuffers& w, uint64_t seed, uint32_t stream, uint64_t first_rollout_id) override {
    bmcts_step_in in{};
    in.n = w.n;
    in.memory = BMCTS_MEM_HOST;
    in.pose = w.pose.data();
    in.alive = w.alive.data();
    in.hyp_id = w.hyp_id.data();
    in.depth = w.depth.data();
    in.action = w.action.data();
    in.draw_id = w.draw_id.data();
    in.seed = seed;
    in.stream = stream;
    bmcts_step_out so{};
    so.next_pose = w.next_pose.data();
    so.next_alive = w.next_alive.data();
    so.reward = w.reward.data();
    so.cost = w.cost.data();
    so.flags = w.flags.data();
    so.last_action = w.last_action.data();
    bmcts_rollout_out ro{};
    ro.result = w.result.data();
    ro.ret_agents = w.ret_agents.data();
    calls_++;
    return oracle_expand_rollout_batch(&cfg_, &scn_, &in, &so, first_rollout_id, &ro, 1);
  }
  int Likelihood(const std::vector<float>& pose, uint32_t alive, const std::vector<float>& action,
                 int n_samples, int n_buckets, float lower, float upper, uint64_t seed,
                 std::vector<float>* prob) override {
    bmcts_likelihood_in in{};
    in.n_worlds = 1;
    in.memory = BMCTS_MEM_HOST;
    in.pose = pose.data();
    in.alive = &alive;
    in.action = action.data();
    in.n_samples = n_samples;
    in.n_buckets = n_buckets;
    in.buckets_lower = lower;
    in.buckets_upper = upper;
    in.seed = seed;
    prob->assign((size_t)scn_.n_agents * scn_.n_hypotheses, 0.0f);
    if (scn_.n_hypotheses == 0) return BMCTS_OK;
    return oracle_likelihood_batch(&cfg_, &scn_, &in, prob->data());
  }
  uint64_t LaunchCount() const override { return 0; }  // no kernels: this is the CPU checker
  std::string LastError() const override { return "oracle evaluator error"; }

 private:
  bmcts_config cfg_;
  bmcts_scenario scn_;
  uint64_t calls_ = 0;
};

}  // namespace

// the product's factory is replaced in THIS library only (liboracle_planner.so)
std::unique_ptr<LeafEvaluator> MakeDefaultEvaluator(int) {
  return std::unique_ptr<LeafEvaluator>(new OracleLeafEvaluator());
}

}  // namespace bark_mcts_b200

extern "C" int oracle_planner_create(int kind, bmcts_planner** out) {
  int st = bmcts_planner_create(kind, 0, out);
  if (st != BMCTS_OK) return st;
  (*out)->evaluator_factory = []() {
    return std::shared_ptr<bark_mcts_b200::LeafEvaluator>(bark_mcts_b200::MakeDefaultEvaluator(0));
  };
  return BMCTS_OK;
}
