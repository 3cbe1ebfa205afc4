// gpu_evaluator.cpp -- LeafEvaluator over the C ABI of the CUDA engine (include/bmcts_c.h).
#include <stdexcept>

#include "evaluator.hpp"

namespace bark_mcts_b200 {

namespace {

class GpuLeafEvaluator : public LeafEvaluator {
 public:
  explicit GpuLeafEvaluator(int device) : device_(device) {}
  ~GpuLeafEvaluator() override {
    if (engine_) bmcts_destroy(engine_);
  }

  int Configure(const bmcts_config& cfg, const bmcts_scenario& scn) override {
    if (!engine_) {
      int st = bmcts_create(device_, &cfg, &engine_);
      if (st != BMCTS_OK) {
        err_ = std::string("bmcts_create: ") + bmcts_status_string(st);
        return st;
      }
      // one launch per MCTS wave and nobody reads the kernel time: no event records
      bmcts_set_kernel_timing(engine_, 0);
    } else {
      int st = bmcts_set_config(engine_, &cfg);
      if (st != BMCTS_OK) return st;
    }
    A_ = scn.n_agents;
    H_ = scn.n_hypotheses;
    return bmcts_set_scenario(engine_, &scn);
  }

  int ExpandRollout(WaveBuffers& w, uint64_t seed, uint32_t stream,
                    uint64_t first_rollout_id) override {
    return Submit(w, seed, stream, first_rollout_id, false);
  }
  int BeginExpandRollout(WaveBuffers& w, uint64_t seed, uint32_t stream,
                         uint64_t first_rollout_id) override {
    return Submit(w, seed, stream, first_rollout_id, true);
  }
  int EndExpandRollout() override { return bmcts_expand_rollout_batch_end(engine_); }

  int Submit(WaveBuffers& w, uint64_t seed, uint32_t stream, uint64_t first_rollout_id,
             bool async) {
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
    return async ? bmcts_expand_rollout_batch_begin(engine_, &in, &so, first_rollout_id, &ro)
                 : bmcts_expand_rollout_batch(engine_, &in, &so, first_rollout_id, &ro);
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
    in.stream = 0;
    prob->assign((size_t)A_ * H_, 0.0f);
    if (H_ == 0) return BMCTS_OK;
    return bmcts_likelihood_batch(engine_, &in, prob->data());
  }

  uint64_t LaunchCount() const override { return bmcts_launch_count(engine_); }
  std::string LastError() const override {
    return engine_ ? std::string(bmcts_last_error(engine_)) : err_;
  }

 private:
  int device_;
  bmcts_engine* engine_ = nullptr;
  int A_ = 0, H_ = 0;
  std::string err_;
};

}  // namespace

std::unique_ptr<LeafEvaluator> MakeDefaultEvaluator(int device) {
  return std::unique_ptr<LeafEvaluator>(new GpuLeafEvaluator(device));
}

}  // namespace bark_mcts_b200
