// evaluator.hpp -- the seam between the host tree search and the rollout engine.
//
// In the reference this seam is the 4th template argument of mcts::Mcts<S, SE, SO, H>
// (behavior_uct_hypothesis.cpp:59-60: H = mcts::RandomHeuristic) together with the state
// concept S::execute / S::plan_action_current_hypothesis / S::get_probability.  Here it is
// an interface over wave-sized batches; the product implementation (gpu_evaluator.cpp)
// forwards to the C ABI of libbmcts.so.  Arrays are agent-major (include/bmcts_c.h).
#pragma once

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "bmcts_c.h"

namespace bark_mcts_b200 {

struct WaveBuffers {
  int n = 0, A = 0;
  // inputs
  std::vector<float> pose;       // [A][n][4]
  std::vector<uint32_t> alive;   // [n]
  std::vector<uint8_t> hyp_id;   // [A][n]
  std::vector<uint16_t> depth;   // [n]
  std::vector<uint8_t> action;   // [A][n]
  std::vector<uint64_t> draw_id; // [A][n]
  // expansion outputs
  std::vector<float> next_pose, reward, cost, last_action;
  std::vector<uint32_t> next_alive, flags;
  // rollout outputs
  std::vector<bmcts_rollout_result> result;
  std::vector<float> ret_agents;

  void resize(int n_, int A_) {
    n = n_;
    A = A_;
    size_t nA = (size_t)n * A;
    pose.assign(nA * 4, 0.0f);
    alive.assign(n, 0);
    hyp_id.assign(nA, 0);
    depth.assign(n, 1);
    action.assign(nA, 0);
    draw_id.assign(nA, 0);
    next_pose.assign(nA * 4, 0.0f);
    reward.assign(nA, 0.0f);
    cost.assign((size_t)2 * n, 0.0f);
    last_action.assign(nA, 0.0f);
    next_alive.assign(n, 0);
    flags.assign(n, 0);
    result.assign(n, bmcts_rollout_result{});
    ret_agents.assign(nA, 0.0f);
  }
  size_t at(int a, int k) const { return (size_t)a * n + k; }
};

class LeafEvaluator {
 public:
  virtual ~LeafEvaluator() {}
  virtual int Configure(const bmcts_config& cfg, const bmcts_scenario& scn) = 0;
  // expansion step with explicit joint actions, then a rollout from every non-terminal child
  virtual int ExpandRollout(WaveBuffers& w, uint64_t seed, uint32_t stream,
                            uint64_t first_rollout_id) = 0;
  // The same in two halves, so that one thread can keep the waves of several trees in flight:
  // Begin enqueues the work and returns, End waits for it and fills the wave's outputs.  The
  // default runs the batch inside Begin.
  virtual int BeginExpandRollout(WaveBuffers& w, uint64_t seed, uint32_t stream,
                                 uint64_t first_rollout_id) {
    deferred_status_ = ExpandRollout(w, seed, stream, first_rollout_id);
    return BMCTS_OK;
  }
  virtual int EndExpandRollout() {
    const int st = deferred_status_;
    deferred_status_ = BMCTS_OK;
    return st;
  }
  // P(action | hypothesis) for every (agent, hypothesis) of one world: prob[a*H + h]
  virtual int Likelihood(const std::vector<float>& pose, uint32_t alive,
                         const std::vector<float>& action, int n_samples, int n_buckets,
                         float lower, float upper, uint64_t seed, std::vector<float>* prob) = 0;
  virtual uint64_t LaunchCount() const = 0;
  virtual std::string LastError() const = 0;

 private:
  int deferred_status_ = BMCTS_OK;
};

// Product factory: the CUDA engine on `device`.  Throws std::runtime_error when the
// engine cannot be created (no GPU) -- there is no CPU fallback in the product.
std::unique_ptr<LeafEvaluator> MakeDefaultEvaluator(int device);

}  // namespace bark_mcts_b200
