// node_prior.hpp -- the neural warm-start seam of the tree search.
//
// In the reference, mcts::NeuralCostConstrainedStatistic asks a learned model for
// {Return, EnvelopeRisk, CollisionRisk, Policy} per ego action the first time a node chooses an
// action (mcts_statistics/mcts_neural_cost_constrained_statistic.hpp:55-59,175-198): the world of
// the node is turned into a vector by a bark-ml observer, the model is evaluated on that ONE
// vector, the policy becomes the node's exploration policy and the values initialise its
// per-action statistics.  Here the same quantities enter through NodePriorProvider, evaluated
// on ALL nodes a wave created in one call, so a model on the GPU sees one batch per wave
// instead of one vector per node.
//
// bark-ml (observer, model loader, value converter) is not in the container: the observer below
// restates NearestObserver from its parameters (tests/data/nheuristic_test.json:132-149); the
// model sits behind a callback.  Parity with the reference is unpinned for this seam.
#pragma once

#include <algorithm>
#include <cmath>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "bmcts_planner_c.h"
#include "params.hpp"

namespace bark_mcts_b200 {

enum { kPriorHasPolicy = 1, kPriorHasValues = 2 };

// ML::NearestObserver: [x, y, theta, v] of the ego, then of the NNearestAgents nearest other
// agents closer than MaxDist (nearest first), every entry min-max normalised to [0, 1]: x and
// y over the bounding box of the road polygon, theta over [MinTheta, MaxTheta], v over
// [MinVel, MaxVel]; missing agents are zero rows.
struct NearestObserver {
  int n_nearest = 4;
  double min_vel = 0.0, max_vel = 50.0, max_dist = 75.0;
  double min_theta = -3.141592653589793, max_theta = 3.141592653589793;
  double min_x = 0.0, max_x = 1.0, min_y = 0.0, max_y = 1.0;

  static NearestObserver FromParams(const Params& p) {
    NearestObserver o;
    o.n_nearest = (int)p.GetInt("ML::NearestObserver::NNearestAgents", 4);
    o.min_vel = p.GetReal("ML::NearestObserver::MinVel", 0.0);
    o.max_vel = p.GetReal("ML::NearestObserver::MaxVel", 50.0);
    o.max_dist = p.GetReal("ML::NearestObserver::MaxDist", 75.0);
    o.min_theta = p.GetReal("ML::NearestObserver::MinTheta", -3.141592653589793);
    o.max_theta = p.GetReal("ML::NearestObserver::MaxTheta", 3.141592653589793);
    return o;
  }
  void SetWorldBounds(const bmcts_scenario& map) {
    if (map.n_road_pts < 1) return;
    min_x = max_x = map.road_x[0];
    min_y = max_y = map.road_y[0];
    for (int i = 1; i < map.n_road_pts; ++i) {
      min_x = std::min<double>(min_x, map.road_x[i]);
      max_x = std::max<double>(max_x, map.road_x[i]);
      min_y = std::min<double>(min_y, map.road_y[i]);
      max_y = std::max<double>(max_y, map.road_y[i]);
    }
  }
  int observation_size() const { return (n_nearest + 1) * 4; }

  static float Norm(double v, double lo, double hi) {
    return hi > lo ? (float)((v - lo) / (hi - lo)) : 0.0f;
  }
  void Row(const float* pose4, float* out) const {
    out[0] = Norm(pose4[0], min_x, max_x);
    out[1] = Norm(pose4[1], min_y, max_y);
    out[2] = Norm(pose4[2], min_theta, max_theta);
    out[3] = Norm(pose4[3], min_vel, max_vel);
  }
  // pose: [A][4] with slot 0 = ego; alive: bit per slot
  void Observe(const float* pose, uint32_t alive, int A, float* out) const {
    std::fill(out, out + observation_size(), 0.0f);
    if (!(alive & 1u)) return;
    Row(pose, out);
    int order[BMCTS_MAX_AGENTS];
    double dist[BMCTS_MAX_AGENTS];
    int n = 0;
    for (int s = 1; s < A; ++s) {
      if (!((alive >> s) & 1u)) continue;
      const double dx = (double)pose[s * 4] - pose[0], dy = (double)pose[s * 4 + 1] - pose[1];
      const double d = std::sqrt(dx * dx + dy * dy);
      if (d >= max_dist) continue;
      dist[s] = d;
      order[n++] = s;
    }
    std::stable_sort(order, order + n, [&](int a, int b) { return dist[a] < dist[b]; });
    for (int i = 0; i < n && i < n_nearest; ++i) Row(pose + order[i] * 4, out + (i + 1) * 4);
  }
};

// values of one node, per ego action
struct NodePrior {
  int mask = 0;  // kPriorHasPolicy | kPriorHasValues
  std::vector<float> policy, ret, envelope_risk, collision_risk;
};

class NodePriorProvider {
 public:
  virtual ~NodePriorProvider() {}
  // observations: [n][obs_dim]; outputs [n][n_actions] each; returns the mask of what was
  // filled (kPriorHas*), or throws
  virtual int Evaluate(int n, int obs_dim, const float* observations, int n_actions,
                       float* policy, float* ret, float* envelope_risk, float* collision_risk) = 0;
};

// a C callback (bmcts_planner_set_node_prior); calls are serialised over the tree threads
class CallbackPriorProvider : public NodePriorProvider {
 public:
  CallbackPriorProvider(bmcts_node_prior_fn fn, void* user) : fn_(fn), user_(user) {}
  int Evaluate(int n, int obs_dim, const float* observations, int n_actions, float* policy,
               float* ret, float* envelope_risk, float* collision_risk) override {
    std::lock_guard<std::mutex> lock(mu_);
    const int mask = fn_(user_, n, obs_dim, observations, n_actions, policy, ret, envelope_risk,
                         collision_risk);
    if (mask < 0) throw std::runtime_error("node prior callback failed (" + std::to_string(mask) + ")");
    calls_++;
    nodes_ += (uint64_t)n;
    return mask;
  }
  uint64_t calls() const { return calls_; }
  uint64_t nodes() const { return nodes_; }

 private:
  bmcts_node_prior_fn fn_;
  void* user_;
  std::mutex mu_;
  uint64_t calls_ = 0, nodes_ = 0;
};

}  // namespace bark_mcts_b200
