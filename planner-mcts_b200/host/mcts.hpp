// mcts.hpp -- host tree search, wave-batched over the rollout engine.
//
// Restates mamcts mcts::Mcts<S, SE, SO, H>::search / iterate and StageNode::select_or_expand
// (SURVEY.md section 3.1 and Appendix B).  The difference to the reference is the seam: an
// iteration does not call S::execute and H::calculate_heuristic_values itself; a WAVE of
// iterations is selected first (virtual visits keep them apart), their expansions +
// rollouts go to the engine as ONE batch (LeafEvaluator::ExpandRollout), and the results
// are backed up in selection order.  WAVE_SIZE = 1 is the strictly sequential search.
#pragma once

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <map>
#include <memory>
#include <random>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "evaluator.hpp"
#include "node_prior.hpp"
#include "statistics.hpp"

namespace bark_mcts_b200 {

enum class EgoMode { kUct, kCostConstrained };
enum class OtherMode { kHypothesis, kUct };

struct RootStats {  // what root-parallel trees / ranks merge (mamcts merge_node_statistics)
  std::vector<double> count, ret_sum, cost0_sum, cost1_sum;
  double iterations = 0, nodes = 0;
  void Resize(int n) {
    count.assign(n, 0.0);
    ret_sum.assign(n, 0.0);
    cost0_sum.assign(n, 0.0);
    cost1_sum.assign(n, 0.0);
  }
  void Add(const RootStats& o) {
    for (size_t i = 0; i < count.size() && i < o.count.size(); ++i) {
      count[i] += o.count[i];
      ret_sum[i] += o.ret_sum[i];
      cost0_sum[i] += o.cost0_sum[i];
      cost1_sum[i] += o.cost1_sum[i];
    }
    iterations += o.iterations;
    nodes += o.nodes;
  }
  std::vector<double> Flatten() const {
    std::vector<double> f;
    for (const auto* v : {&count, &ret_sum, &cost0_sum, &cost1_sum}) f.insert(f.end(), v->begin(), v->end());
    f.push_back(iterations);
    f.push_back(nodes);
    return f;
  }
  void Unflatten(const std::vector<double>& f) {
    const size_t n = count.size();
    for (size_t i = 0; i < n; ++i) {
      count[i] = f[i];
      ret_sum[i] = f[n + i];
      cost0_sum[i] = f[2 * n + i];
      cost1_sum[i] = f[3 * n + i];
    }
    iterations = f[4 * n];
    nodes = f[4 * n + 1];
  }
};

// joint action of one expansion: ActionKey per agent slot, fixed capacity (no allocation on
// look-ups); compared over the first `n` slots
struct JointKey {
  int n = 0;
  ActionKey k[BMCTS_MAX_AGENTS];
  explicit JointKey(int n_ = 0) : n(n_) {
    for (int i = 0; i < n; ++i) k[i] = 0;
  }
  ActionKey& operator[](int i) { return k[i]; }
  const ActionKey& operator[](int i) const { return k[i]; }
  bool operator==(const JointKey& o) const {
    if (n != o.n) return false;
    for (int i = 0; i < n; ++i)
      if (k[i] != o.k[i]) return false;
    return true;
  }
};

// 128-bit fingerprint of a joint action (two independent 64-bit mixes over the agents' keys)
struct JointHash {
  uint64_t h1 = 0, h2 = 0;
  explicit JointHash(const JointKey& j) {
    uint64_t a = 0x9E3779B97F4A7C15ull, b = 0xC2B2AE3D27D4EB4Full;
    for (int i = 0; i < j.n; ++i) {
      a = (a ^ j.k[i]) * 0xFF51AFD7ED558CCDull;
      a ^= a >> 29;
      b = (b + j.k[i] + ((uint64_t)i << 32)) * 0x9FB21C651E98DF25ull;
      b ^= b >> 31;
    }
    h1 = a ^ (a >> 32);
    h2 = b ^ (b >> 33);
  }
};

struct Node;

// Children of a node by joint action: open addressing over the fingerprints, 24 bytes per slot,
// one cache line touched per look-up.  (A node-based map with the 132-byte joint action as key
// cost 2300 cycles per insertion at the root of a 32-agent search.)  Two joint actions with the
// same 128-bit fingerprint would share a child: with 10^6 children the odds are below 10^-26.
class ChildTable {
 public:
  Node* Find(const JointHash& h) const {
    if (slots_.empty()) return nullptr;
    const size_t mask = slots_.size() - 1;
    for (size_t i = h.h1 & mask;; i = (i + 1) & mask) {
      const Slot& s = slots_[i];
      if (!s.node) return nullptr;
      if (s.h1 == h.h1 && s.h2 == h.h2) return s.node;
    }
  }
  // the child of this joint action, or the (empty) place where the caller puts the new child
  Node** FindOrClaim(const JointHash& h) {
    if ((used_ + 1) * 2 > slots_.size()) Grow();
    const size_t mask = slots_.size() - 1;
    for (size_t i = h.h1 & mask;; i = (i + 1) & mask) {
      Slot& s = slots_[i];
      if (!s.node) {
        s.h1 = h.h1;
        s.h2 = h.h2;
        used_++;  // the caller stores the node right away
        return &s.node;
      }
      if (s.h1 == h.h1 && s.h2 == h.h2) return &s.node;
    }
  }
  size_t size() const { return used_; }

 private:
  struct Slot {
    uint64_t h1 = 0, h2 = 0;
    Node* node = nullptr;
  };
  void Grow() {
    std::vector<Slot> old;
    old.swap(slots_);
    slots_.assign(old.empty() ? 8 : old.size() * 2, Slot());
    const size_t mask = slots_.size() - 1;
    for (const Slot& s : old) {
      if (!s.node) continue;
      size_t i = s.h1 & mask;
      while (slots_[i].node) i = (i + 1) & mask;
      slots_[i] = s;
    }
  }
  std::vector<Slot> slots_;
  size_t used_ = 0;
};

struct Node {
  Node* parent = nullptr;
  unsigned state_depth = 1;  // MctsStateBase::depth_
  int tree_depth = 0;
  float* pose = nullptr;     // [A][4], in the search's pose arena
  uint32_t alive = 0;
  bool terminal = false;
  bool stats_ready = false;  // action tables built (first time the node is selected through)
  // The ego statistic of the search's mode (UctStatistic or CostConstrainedStatistic) exists
  // from the first time the node is selected through.  Until then -- for most nodes: for ever --
  // the node only carries its heuristic estimate: the last value and how often it was set.
  std::unique_ptr<UctStatistic> ego_uct;
  std::unique_ptr<CostConstrainedStatistic> ego_cc;
  double h_return = 0.0, h_cost[2] = {0.0, 0.0};
  unsigned h_visits = 0;
  double LatestReturn() const {
    return ego_uct ? ego_uct->latest_return() : ego_cc ? ego_cc->latest_return() : h_return;
  }
  double LatestCost(int i) const { return ego_cc ? ego_cc->latest_cost(i) : h_cost[i]; }
  std::vector<HypothesisStatistic> other_hyp;  // per agent slot (slot 0 unused); built lazily
  std::vector<UctStatistic> other_uct;
  // latest back-propagated value seen by the other agents' statistics of this node: the ego
  // cost (hypothesis states: the same for every agent) / the agent's own return (cooperative)
  double latest_cost_other = 0.0;
  std::vector<double> latest_ret_other;
  ChildTable children;
  // rewards[] of the execute() that created this node: the ego's, and (cooperative states
  // only: nobody else reads them) every agent's
  double edge_reward_ego = 0.0;
  std::vector<double> edge_reward;
  double edge_cost[2] = {0.0, 0.0};
  std::unique_ptr<NodePrior> prior;  // neural warm start of the ego statistic (node_prior.hpp)
};

class Mcts {
 public:
  Mcts(const MctsParameters& params, EgoMode ego_mode, OtherMode other_mode, int num_agents,
       int num_ego_actions, int num_hypotheses, int num_costs, LeafEvaluator* evaluator,
       unsigned tree_index = 0)
      : p_(params), ego_mode_(ego_mode), other_mode_(other_mode), A_(num_agents),
        n_actions_(num_ego_actions), H_(num_hypotheses), n_costs_(num_costs), eval_(evaluator),
        tree_index_(tree_index), gen_(params.RANDOM_SEED + tree_index),
        lambdas_(params.cost_constrained_statistic.LAMBDAS) {
    lambdas_.resize(std::max(1, n_costs_), lambdas_.empty() ? 1.0 : lambdas_.back());
  }

  // hypothesis sampler: fills hyp[slot] for slots 1..A-1 (root-belief sampling); may be null
  using HypothesisSampler = std::function<void(std::vector<uint8_t>*)>;

  void Search(const std::vector<float>& root_pose, uint32_t root_alive, unsigned root_depth,
              const HypothesisSampler& sample_hypotheses) {
    StartSearch(root_pose, root_alive, root_depth, sample_hypotheses);
    while (!Done()) {
      BeginWave();
      EndWave();
    }
    FinishSearch();
  }

  // The search in steps, for a driver that keeps the waves of several trees in flight from one
  // thread (BehaviorUCTBase::RunTrees): StartSearch, then BeginWave / EndWave until Done(),
  // then FinishSearch.  BeginWave selects a wave and hands it to the engine without waiting;
  // EndWave waits for the engine and backs the results up.
  void StartSearch(const std::vector<float>& root_pose, uint32_t root_alive, unsigned root_depth,
                   const HypothesisSampler& sample_hypotheses) {
    t_start_ = std::chrono::steady_clock::now();
    sampler_ = sample_hypotheses;
    node_chunks_.clear();
    pose_chunks_.clear();
    n_nodes_ = 0;
    awaiting_prior_.clear();
    cc_params_ = p_;
    cc_params_.cost_constrained_statistic.COST_CONSTRAINTS = constraints_;
    root_ = NewNode(nullptr, &root_pose, root_alive, root_depth, false);
    EvaluatePriors();
    EnsureStats(root_);
    t_select_ = t_fill_ = t_engine_ = t_backup_ = 0.0;
    iterations_ = 0;
    max_tree_depth_ = 0;
    wave_in_flight_ = 0;
    wave_size_ = p_.WAVE_SIZE > 0
                     ? p_.WAVE_SIZE
                     : std::min<long>(256, std::max<long>(1, p_.MAX_NUMBER_OF_ITERATIONS / 128));
  }

  bool Done() const {
    if (iterations_ >= p_.MAX_NUMBER_OF_ITERATIONS) return true;
    const double ms =
        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start_).count();
    return iterations_ > 0 && ms > (double)p_.MAX_SEARCH_TIME;
  }

  void FinishSearch() {
    search_time_ms_ =
        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start_).count();
    if (std::getenv("BMCTS_DEBUG"))
      std::fprintf(stderr, "[bmcts] search %.2f ms: select %.2f, fill %.2f, engine %.2f, backup %.2f (%ld iterations, %zu nodes)\n",
                   search_time_ms_, t_select_, t_fill_, t_engine_, t_backup_, iterations_, n_nodes_);
  }

  RootStats GetRootStats() const {
    RootStats s;
    s.Resize(n_actions_);
    const UctStatistic& r = ego_mode_ == EgoMode::kUct ? *root_->ego_uct : root_->ego_cc->reward_statistic();
    for (int a = 0; a < n_actions_; ++a) {
      s.count[a] = r.pairs()[a].count;
      s.ret_sum[a] = r.pairs()[a].value * r.pairs()[a].count;
      if (ego_mode_ == EgoMode::kCostConstrained) {
        s.cost0_sum[a] = root_->ego_cc->cost_statistic(0).pairs()[a].value * r.pairs()[a].count;
        if (n_costs_ > 1)
          s.cost1_sum[a] = root_->ego_cc->cost_statistic(1).pairs()[a].value * r.pairs()[a].count;
      }
    }
    s.iterations = (double)iterations_;
    s.nodes = (double)n_nodes_;
    return s;
  }

  long numIterations() const { return iterations_; }
  size_t numNodes() const { return n_nodes_; }
  double searchTime() const { return search_time_ms_; }
  int maxTreeDepth() const { return max_tree_depth_; }
  const Node& get_root() const { return *root_; }
  const std::vector<double>& lambdas() const { return lambdas_; }
  void set_cost_constraints(const std::vector<double>& c) { constraints_ = c; }
  // neural warm start (cost-constrained ego statistic only): every node a wave creates is
  // observed and the provider is asked once per wave for the whole batch
  void set_node_prior(NodePriorProvider* provider, const NearestObserver& observer,
                      double prior_visits) {
    prior_ = provider;
    observer_ = observer;
    prior_visits_ = prior_visits;
  }
  // mamcts MctsEdgeInfo / MctsStateInfo (behavior_uct_base.hpp:26-27,63-67: what the reference
  // extracts for its tree viewer and its tree-shape tests): one record per (node, agent slot,
  // expanded action) up to max_depth -- depth of the node, action (ego: macro-action index;
  // others under hypotheses: bit pattern of the sampled acceleration), share of the node's visits
  struct EdgeInfo {
    int slot = 0;
    unsigned depth = 0;
    uint32_t action = 0;
    double weight = 0.0;
  };
  std::vector<EdgeInfo> ExtractEdges(unsigned max_depth) const {
    std::vector<EdgeInfo> out;
    ForEachNode([&](const Node& n) {
      if (!n.stats_ready || (unsigned)n.tree_depth > max_depth) return;
      const UctStatistic* ego = n.ego_uct ? n.ego_uct.get() : n.ego_cc ? &n.ego_cc->reward_statistic() : nullptr;
      if (ego) {
        double tot = 0.0;
        for (const auto& pr : ego->pairs()) tot += pr.count;
        for (int a = 0; a < (int)ego->pairs().size(); ++a)
          if (ego->pairs()[a].count > 0)
            out.push_back({0, (unsigned)n.tree_depth, (uint32_t)a, ego->pairs()[a].count / std::max(1.0, tot)});
      }
      for (int s = 1; s < A_; ++s) {
        if (other_mode_ == OtherMode::kHypothesis) {
          if (n.other_hyp.empty()) continue;
          std::map<ActionKey, double> by_key;
          double tot = 0.0;
          for (int h = 0; h < n.other_hyp[s].num_hypotheses(); ++h)
            for (const auto& e : n.other_hyp[s].table(h))
              if (e.resolved && e.count > 0) {
                by_key[e.key] += e.count;
                tot += e.count;
              }
          for (const auto& kv : by_key) out.push_back({s, (unsigned)n.tree_depth, kv.first, kv.second / std::max(1.0, tot)});
        } else if (!n.other_uct.empty()) {
          double tot = 0.0;
          for (const auto& pr : n.other_uct[s].pairs()) tot += pr.count;
          for (int a = 0; a < (int)n.other_uct[s].pairs().size(); ++a)
            if (n.other_uct[s].pairs()[a].count > 0)
              out.push_back({s, (unsigned)n.tree_depth, (uint32_t)a, n.other_uct[s].pairs()[a].count / std::max(1.0, tot)});
        }
      }
    });
    return out;
  }
  struct StateInfo {
    unsigned depth = 0;
    uint32_t alive = 0;
    std::vector<float> pose;  // [A][4]
  };
  std::vector<StateInfo> ExtractStates(unsigned max_depth) const {
    std::vector<StateInfo> out;
    ForEachNode([&](const Node& n) {
      if ((unsigned)n.tree_depth > max_depth) return;
      out.push_back({(unsigned)n.tree_depth, n.alive, std::vector<float>(n.pose, n.pose + (size_t)A_ * 4)});
    });
    return out;
  }

  // number of distinct actions expanded at the root per agent slot (tree-shape tests)
  std::vector<int> RootExpandedActions() const {
    std::vector<int> n(A_, 0);
    for (int a = 0; a < n_actions_; ++a) {
      const UctStatistic& r = ego_mode_ == EgoMode::kUct ? *root_->ego_uct : root_->ego_cc->reward_statistic();
      n[0] += r.pairs()[a].count > 0;
    }
    for (int s = 1; s < A_; ++s) {
      if (other_mode_ == OtherMode::kHypothesis) {
        n[s] = root_->other_hyp[s].NumExpandedActions();
      } else {
        for (const auto& pr : root_->other_uct[s].pairs()) n[s] += pr.count > 0;
      }
    }
    return n;
  }

 private:
  struct PathStep {
    Node* node = nullptr;
    int ego_action = 0;
    int other_idx[BMCTS_MAX_AGENTS];      // HypothesisStatistic entry index / UCT action per slot
    uint8_t other_new[BMCTS_MAX_AGENTS];  // the entry's action is not known yet (plan it)
  };
  struct Selection {
    std::vector<PathStep> path;  // path.back() is the expansion step (unless no_expand)
    std::vector<uint8_t> hyp;    // sampled hypothesis per slot for this iteration
    bool needs_engine = false;
    bool rollout_only = false;   // reached a non-terminal node at MAX_SEARCH_DEPTH: heuristic only
    Node* reached = nullptr;     // existing terminal / max-depth node reached instead
  };

  Node* NewNode(Node* parent, const std::vector<float>* pose, uint32_t alive, unsigned state_depth,
                bool terminal) {
    // nodes and their poses come from chunked arenas: one allocation per 256 nodes
    const size_t slot = n_nodes_ % kChunk;
    if (slot == 0) {
      node_chunks_.emplace_back(new Node[kChunk]);
      pose_chunks_.emplace_back(new float[kChunk * (size_t)A_ * 4]);
    }
    Node* n = &node_chunks_.back()[slot];
    n->pose = &pose_chunks_.back()[slot * (size_t)A_ * 4];
    n_nodes_++;
    n->parent = parent;
    n->tree_depth = parent ? parent->tree_depth + 1 : 0;
    max_tree_depth_ = std::max(max_tree_depth_, n->tree_depth);
    n->state_depth = state_depth;
    if (pose) std::copy(pose->begin(), pose->begin() + (size_t)A_ * 4, n->pose);  // else: the caller fills it
    n->alive = alive;
    n->terminal = terminal;
    if (other_mode_ == OtherMode::kUct) n->edge_reward.assign(A_, 0.0);
    if (other_mode_ == OtherMode::kUct) n->latest_ret_other.assign(A_, 0.0);
    if (prior_ && !terminal && ego_mode_ == EgoMode::kCostConstrained) awaiting_prior_.push_back(n);
    return n;
  }

  // the one node that stands in for children created beyond MAX_NUMBER_OF_NODES
  Node* ScratchNode(Node* parent, uint32_t alive, unsigned state_depth, bool terminal) {
    if (!scratch_) {
      scratch_.reset(new Node());
      scratch_pose_.reset(new float[(size_t)A_ * 4]);
    }
    *scratch_ = Node();
    Node* n = scratch_.get();
    n->pose = scratch_pose_.get();
    n->parent = parent;
    n->tree_depth = parent ? parent->tree_depth + 1 : 0;
    n->state_depth = state_depth;
    n->alive = alive;
    n->terminal = terminal;
    if (other_mode_ == OtherMode::kUct) n->edge_reward.assign(A_, 0.0);
    if (other_mode_ == OtherMode::kUct) n->latest_ret_other.assign(A_, 0.0);
    return n;
  }

  void EnsureStats(Node* n) {
    if (n->stats_ready) return;
    n->stats_ready = true;
    if (ego_mode_ == EgoMode::kUct) {
      n->ego_uct = std::make_unique<UctStatistic>();
      n->ego_uct->Init(n_actions_, p_.uct_statistic.LOWER_BOUND, p_.uct_statistic.UPPER_BOUND,
                       p_.uct_statistic.EXPLORATION_CONSTANT, p_.DISCOUNT_FACTOR,
                       p_.USE_BOUND_ESTIMATION, p_.uct_statistic.PROGRESSIVE_WIDENING_K,
                       p_.uct_statistic.PROGRESSIVE_WIDENING_ALPHA);
      n->ego_uct->SeedHeuristic(n->h_return, n->h_visits);
    } else {
      n->ego_cc = std::make_unique<CostConstrainedStatistic>();
      n->ego_cc->Init(n_actions_, n_costs_, cc_params_);
      n->ego_cc->SeedHeuristic(n->h_return, n->h_cost, n->h_visits);
      if (n->prior) {
        const NodePrior& pr = *n->prior;
        const bool vals = (pr.mask & kPriorHasValues) != 0;
        n->ego_cc->SetPrior((pr.mask & kPriorHasPolicy) ? pr.policy.data() : nullptr,
                           vals ? pr.ret.data() : nullptr, vals ? pr.envelope_risk.data() : nullptr,
                           vals ? pr.collision_risk.data() : nullptr, prior_visits_);
      }
    }
    if (other_mode_ == OtherMode::kHypothesis) {
      n->other_hyp.assign(A_, HypothesisStatistic());
      for (int s = 1; s < A_; ++s) n->other_hyp[s].Init(H_, p_);
    } else {
      n->other_uct.assign(A_, UctStatistic());
      for (int s = 1; s < A_; ++s)
        n->other_uct[s].Init(n_actions_, p_.uct_statistic.LOWER_BOUND, p_.uct_statistic.UPPER_BOUND,
                             p_.uct_statistic.EXPLORATION_CONSTANT, p_.DISCOUNT_FACTOR,
                             p_.USE_BOUND_ESTIMATION, p_.uct_statistic.PROGRESSIVE_WIDENING_K,
                             p_.uct_statistic.PROGRESSIVE_WIDENING_ALPHA);
    }
  }

  // StageNode::select_or_expand down the tree for one iteration, from `node` on
  // (fills a Selection that is reused from wave to wave: no allocation per iteration)
  void SelectFrom(Node* node, Selection& sel) {
    for (;;) {
      if (node->terminal || node->tree_depth >= p_.MAX_SEARCH_DEPTH) {
        sel.reached = node;
        // mamcts iterate(): the heuristic also runs on a non-terminal node at the depth cap
        sel.rollout_only = sel.needs_engine = !node->terminal;
        return;
      }
      EnsureStats(node);
      PathStep st;
      st.node = node;
      for (int s = 0; s < A_; ++s) {
        st.other_idx[s] = -1;
        st.other_new[s] = 0;
      }
      st.ego_action = ego_mode_ == EgoMode::kUct ? node->ego_uct->ChooseNextAction(gen_)
                                                 : node->ego_cc->ChooseNextAction(lambdas_, gen_);
      for (int s = 1; s < A_; ++s) {
        if (!((node->alive >> s) & 1u)) continue;
        ChooseOther(node, s, sel, st);
      }
      sel.path.push_back(st);
      if (Node* child = ResolvedChild(node, sel, st)) {
        node = child;
        continue;
      }
      sel.needs_engine = true;
      return;
    }
  }

  // another agent's choice at `node` for this iteration (HypothesisStatistic / UctStatistic)
  void ChooseOther(Node* node, int s, const Selection& sel, PathStep& st) {
    if (other_mode_ == OtherMode::kHypothesis) {
      bool is_new = false;
      st.other_idx[s] = node->other_hyp[s].ChooseNextAction(sel.hyp[s], gen_, next_draw_id_++, &is_new);
      st.other_new[s] = is_new;
    } else {
      st.other_idx[s] = node->other_uct[s].ChooseNextAction(gen_);
    }
  }

  // the child of the chosen joint action if every agent's action is known and it was expanded
  Node* ResolvedChild(const Node* node, const Selection& sel, const PathStep& st) const {
    JointKey joint(A_);
    joint[0] = (ActionKey)st.ego_action;
    for (int s = 1; s < A_; ++s) {
      if (st.other_idx[s] < 0) {
        joint[s] = ~ActionKey(0);
      } else if (other_mode_ == OtherMode::kHypothesis) {
        if (st.other_new[s]) return nullptr;
        joint[s] = node->other_hyp[s].KeyOf(sel.hyp[s], st.other_idx[s]);
      } else {
        joint[s] = (ActionKey)st.other_idx[s];
      }
    }
    return node->children.Find(JointHash(joint));
  }

  // The selections of one wave.  Every one of them starts at the root, and in the flat trees of
  // many-agent worlds most end there, so the root level is done AGENT BY AGENT over the whole
  // wave: the 64 action tables of one agent stay in the cache while the wave's iterations choose
  // from them, instead of every iteration walking the tables of all agents (a 2 MB working set
  // at 32 agents x 64 hypotheses: one cache miss per table).  Each statistic sees its calls in
  // iteration order, exactly as if the iterations had been selected one after the other; only
  // the interleaving of the shared random generator differs.  A wave of one iteration is the
  // sequential search.
  void SelectWave(int w) {
    root_w_ = 0;  // set when the root level goes through the per-agent arrays
    for (int i = 0; i < w; ++i) {
      Selection& sel = sels_[i];
      sel.path.clear();
      sel.needs_engine = false;
      sel.rollout_only = false;
      sel.reached = nullptr;
      sel.hyp.assign(A_, 0);
      if (sampler_) sampler_(&sel.hyp);
    }
    if (w == 1 || root_->terminal || root_->tree_depth >= p_.MAX_SEARCH_DEPTH) {
      for (int i = 0; i < w; ++i) SelectFrom(root_, sels_[i]);
      return;
    }
    EnsureStats(root_);
    for (int i = 0; i < w; ++i) {
      PathStep st;
      st.node = root_;
      for (int s = 0; s < A_; ++s) {
        st.other_idx[s] = -1;
        st.other_new[s] = 0;
      }
      st.ego_action = ego_mode_ == EgoMode::kUct ? root_->ego_uct->ChooseNextAction(gen_)
                                                 : root_->ego_cc->ChooseNextAction(lambdas_, gen_);
      sels_[i].path.push_back(st);
    }
    if (other_mode_ == OtherMode::kHypothesis) {
      // the wave's root-level choices as arrays per agent ([slot][iteration]): the inner loop
      // streams through them instead of touching a Selection, its path and its hypothesis
      // vector (three cache lines of three different heap blocks) per (agent, iteration)
      root_w_ = w;
      root_hyp_.resize((size_t)A_ * w);
      root_idx_.assign((size_t)A_ * w, -1);
      root_new_.assign((size_t)A_ * w, 0);
      for (int i = 0; i < w; ++i)
        for (int s = 1; s < A_; ++s) root_hyp_[(size_t)s * w + i] = sels_[i].hyp[s];
      for (int s = 1; s < A_; ++s) {
        if (!((root_->alive >> s) & 1u)) continue;
        HypothesisStatistic& hs = root_->other_hyp[s];
        const uint8_t* hyp = &root_hyp_[(size_t)s * w];
        int* idx = &root_idx_[(size_t)s * w];
        uint8_t* isnew = &root_new_[(size_t)s * w];
        for (int i = 0; i < w; ++i) {
          bool n = false;
          idx[i] = hs.ChooseNextAction(hyp[i], gen_, next_draw_id_++, &n);
          isnew[i] = n;
        }
      }
      for (int i = 0; i < w; ++i) {
        PathStep& st = sels_[i].path[0];
        for (int s = 1; s < A_; ++s) {
          st.other_idx[s] = root_idx_[(size_t)s * w + i];
          st.other_new[s] = root_new_[(size_t)s * w + i];
        }
      }
    } else {
      for (int s = 1; s < A_; ++s) {
        if (!((root_->alive >> s) & 1u)) continue;
        for (int i = 0; i < w; ++i) ChooseOther(root_, s, sels_[i], sels_[i].path[0]);
      }
    }
    for (int i = 0; i < w; ++i) {
      Selection& sel = sels_[i];
      if (Node* child = ResolvedChild(root_, sel, sel.path[0]))
        SelectFrom(child, sel);
      else
        sel.needs_engine = true;
    }
  }

  using clock = std::chrono::steady_clock;
  static double Ms(clock::time_point a, clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count();
  }

 public:
  void BeginWave() {
    const int w = (int)std::min<long>(wave_size_, p_.MAX_NUMBER_OF_ITERATIONS - iterations_);
    auto ms = Ms;
    const auto t0 = clock::now();
    std::vector<Selection>& sels = sels_;
    if ((int)sels.size() < w) sels.resize(w);
    SelectWave(w);
    const auto t1 = clock::now();
    t_select_ += ms(t0, t1);
    std::vector<int>& slot = slot_;
    slot.assign(w, -1);
    int n_engine = 0;
    for (int i = 0; i < w; ++i)
      if (sels[i].needs_engine) slot[i] = n_engine++;
    if (n_engine > 0) {
      wave_.resize(n_engine, A_);
      for (int i = 0; i < w; ++i) {
        if (slot[i] < 0) continue;
        const int k = slot[i];
        if (sels[i].rollout_only) {  // no expansion step: BMCTS_ACTION_NONE in the ego's slot
          const Node* nd = sels[i].reached;
          wave_.alive[k] = nd->alive;
          wave_.depth[k] = (uint16_t)std::min<unsigned>(nd->state_depth, BMCTS_MAX_DEPTH - 1);
          for (int a = 0; a < A_; ++a) {
            const size_t e = wave_.at(a, k);
            for (int f = 0; f < 4; ++f) wave_.pose[e * 4 + f] = nd->pose[(size_t)a * 4 + f];
            wave_.hyp_id[e] = sels[i].hyp[a];
            wave_.action[e] = a == 0 ? (uint8_t)BMCTS_ACTION_NONE : (uint8_t)0;
          }
          continue;
        }
        const PathStep& st = sels[i].path.back();
        const Node* nd = st.node;
        wave_.alive[k] = nd->alive;
        wave_.depth[k] = (uint16_t)std::min<unsigned>(nd->state_depth, BMCTS_MAX_DEPTH - 1);
        for (int a = 0; a < A_; ++a) {
          const size_t e = wave_.at(a, k);
          for (int f = 0; f < 4; ++f) wave_.pose[e * 4 + f] = nd->pose[(size_t)a * 4 + f];
          wave_.hyp_id[e] = sels[i].hyp[a];
          if (a == 0) {
            wave_.action[e] = (uint8_t)st.ego_action;
          } else if (st.other_idx[a] >= 0) {
            if (other_mode_ == OtherMode::kHypothesis)
              wave_.draw_id[e] = nd->other_hyp[a].table(sels[i].hyp[a])[st.other_idx[a]].draw_id;
            else
              wave_.action[e] = (uint8_t)st.other_idx[a];
          }
        }
      }
      const auto t2 = clock::now();
      t_fill_ += ms(t1, t2);
      const uint64_t first_id = rollout_counter_;
      rollout_counter_ += (uint64_t)n_engine;
      int stc = eval_->BeginExpandRollout(wave_, p_.RANDOM_SEED, tree_index_, first_id);
      if (stc != BMCTS_OK)
        throw std::runtime_error("rollout engine failed: " + eval_->LastError());
      t_engine_ += ms(t2, clock::now());
    }
    wave_in_flight_ = w;
    wave_engine_ = n_engine > 0;
  }

  void EndWave() {
    const int w = wave_in_flight_;
    wave_in_flight_ = 0;
    std::vector<Selection>& sels = sels_;
    std::vector<int>& slot = slot_;
    auto ms = Ms;
    if (wave_engine_) {
      const auto t2 = clock::now();
      int stc = eval_->EndExpandRollout();
      if (stc != BMCTS_OK)
        throw std::runtime_error("rollout engine failed: " + eval_->LastError());
      t_engine_ += ms(t2, clock::now());
    }
    const auto t3 = clock::now();
    // As in SelectWave, the root level of the wave's back-ups is done agent by agent (the tables
    // of one agent stay in the cache): first every newly planned action of a root expansion is
    // resolved, then the iterations are backed up in order with the root's other-agent updates
    // deferred, then those are applied per agent in iteration order.  Each statistic sees the
    // calls of the sequential order; nothing is selected in between.
    const bool by_agent = w > 1 && other_mode_ == OtherMode::kHypothesis && root_->stats_ready &&
                          root_w_ == w;
    if (by_agent) {
      root_expand_.resize(w);
      for (int i = 0; i < w; ++i)
        root_expand_[i] = sels[i].needs_engine && !sels[i].rollout_only && sels[i].path.size() == 1;
      for (int s = 1; s < A_; ++s) {
        HypothesisStatistic& hs = root_->other_hyp[s];
        const uint8_t* hyp = &root_hyp_[(size_t)s * w];
        int* idx = &root_idx_[(size_t)s * w];
        uint8_t* isnew = &root_new_[(size_t)s * w];
        const float* last = &wave_.last_action[wave_.n > 0 ? wave_.at(s, 0) : 0];
        for (int i = 0; i < w; ++i) {
          if (!root_expand_[i] || idx[i] < 0 || !isnew[i]) continue;
          idx[i] = hs.Resolve(hyp[i], idx[i], KeyOfAcceleration(last[slot[i]]));
          isnew[i] = 0;
        }
      }
      for (int i = 0; i < w; ++i) {
        if (!root_expand_[i]) continue;
        PathStep& st = sels[i].path[0];
        for (int s = 1; s < A_; ++s) {
          st.other_idx[s] = root_idx_[(size_t)s * w + i];
          st.other_new[s] = root_new_[(size_t)s * w + i];
        }
      }
      defer_cost_.assign(w, 0.0);
      defer_latest_.assign(w, 0.0);
    }
    for (int i = 0; i < w; ++i) {
      Backup(sels[i], slot[i], by_agent ? i : -1);
      iterations_++;
      if (ego_mode_ == EgoMode::kCostConstrained) UpdateLambdas();
    }
    if (by_agent) {
      for (int s = 1; s < A_; ++s) {
        HypothesisStatistic& hs = root_->other_hyp[s];
        const uint8_t* hyp = &root_hyp_[(size_t)s * w];
        const int* idx = &root_idx_[(size_t)s * w];
        for (int i = 0; i < w; ++i)
          if (idx[i] >= 0) hs.BackupEntry(hyp[i], idx[i], defer_cost_[i], defer_latest_[i]);
      }
    }
    t_backup_ += ms(t3, clock::now());
    EvaluatePriors();
  }

 private:

  // one provider call for all nodes created since the last call
  void EvaluatePriors() {
    if (!prior_ || awaiting_prior_.empty()) return;
    const int n = (int)awaiting_prior_.size(), d = observer_.observation_size();
    obs_.assign((size_t)n * d, 0.0f);
    for (int i = 0; i < n; ++i)
      observer_.Observe(awaiting_prior_[i]->pose, awaiting_prior_[i]->alive, A_,
                        obs_.data() + (size_t)i * d);
    for (auto* v : {&pr_policy_, &pr_ret_, &pr_env_, &pr_col_}) v->assign((size_t)n * n_actions_, 0.0f);
    const int mask = prior_->Evaluate(n, d, obs_.data(), n_actions_, pr_policy_.data(),
                                      pr_ret_.data(), pr_env_.data(), pr_col_.data());
    for (int i = 0; i < n; ++i) {
      auto pr = std::make_unique<NodePrior>();
      pr->mask = mask;
      const size_t o = (size_t)i * n_actions_;
      if (mask & kPriorHasPolicy) pr->policy.assign(pr_policy_.begin() + o, pr_policy_.begin() + o + n_actions_);
      if (mask & kPriorHasValues) {
        pr->ret.assign(pr_ret_.begin() + o, pr_ret_.begin() + o + n_actions_);
        pr->envelope_risk.assign(pr_env_.begin() + o, pr_env_.begin() + o + n_actions_);
        pr->collision_risk.assign(pr_col_.begin() + o, pr_col_.begin() + o + n_actions_);
      }
      awaiting_prior_[i]->prior = std::move(pr);
    }
    awaiting_prior_.clear();
  }

  void SetHeuristic(Node* n, double ret, const double* cost, const float* ret_agents, int k) {
    SetEgoHeuristic(n, ret, cost);
    if (other_mode_ == OtherMode::kHypothesis) {
      n->latest_cost_other = cost[0];
    } else {
      for (int s = 1; s < A_; ++s)
        n->latest_ret_other[s] = ret_agents ? (double)ret_agents[wave_.at(s, k)] : 0.0;
    }
  }

  // defer >= 0: the other agents' updates at the ROOT are left to the caller (EndWave applies them
  // agent by agent); what they need is stored under this index
  void Backup(Selection& sel, int k, int defer = -1) {
    Node* leaf = sel.reached;
    if (sel.rollout_only) {
      // RandomHeuristic on the depth-capped node itself (node->update_statistics(heuristic))
      const bmcts_rollout_result& r = wave_.result[k];
      const double cost[2] = {r.cost0, r.cost1};
      SetHeuristic(leaf, r.ret_ego, cost, wave_.ret_agents.data(), k);
    } else if (sel.needs_engine) {
      PathStep& st = sel.path.back();
      Node* parent = st.node;
      // resolve the others' newly planned actions and build the joint action key
      JointKey joint(A_);
      joint[0] = (ActionKey)st.ego_action;
      for (int s = 1; s < A_; ++s) {
        if (st.other_idx[s] < 0) {
          joint[s] = ~ActionKey(0);
          continue;
        }
        if (other_mode_ == OtherMode::kHypothesis) {
          const int h = sel.hyp[s];
          if (st.other_new[s]) {
            ActionKey key = KeyOfAcceleration(wave_.last_action[wave_.at(s, k)]);
            st.other_idx[s] = parent->other_hyp[s].Resolve(h, st.other_idx[s], key);
          }
          joint[s] = parent->other_hyp[s].KeyOf(h, st.other_idx[s]);
        } else {
          joint[s] = (ActionKey)st.other_idx[s];
        }
      }
      // one hash of the joint action: look up, and claim the slot if it is new
      const bool link = (long)n_nodes_ < p_.MAX_NUMBER_OF_NODES;
      const JointHash jh(joint);
      Node* existing = nullptr;
      Node** place = nullptr;
      if (link) {
        place = parent->children.FindOrClaim(jh);
        existing = *place;
      } else {
        existing = parent->children.Find(jh);
      }
      const bmcts_rollout_result& r = wave_.result[k];
      const double cost[2] = {r.cost0, r.cost1};
      if (existing) {
        // The newly planned joint action is one that was expanded before (in this wave, or its
        // actions resolved to known ones): the engine has rolled out from that same child state.
        // mamcts would descend through the child; here the fresh rollout is the sample the
        // parents back up, and a child that already has action statistics keeps its own value
        // estimate and visit count.
        leaf = existing;
        if (leaf->stats_ready) {
          if (leaf->ego_uct) leaf->ego_uct->SetLatestReturn(r.ret_ego);
          if (leaf->ego_cc) leaf->ego_cc->SetLatest(r.ret_ego, cost);
          if (other_mode_ == OtherMode::kHypothesis) {
            leaf->latest_cost_other = cost[0];
          } else {
            for (int s = 1; s < A_; ++s)
              leaf->latest_ret_other[s] = (double)wave_.ret_agents[wave_.at(s, k)];
          }
        } else {
          SetHeuristic(leaf, r.ret_ego, cost, wave_.ret_agents.data(), k);
        }
      } else {
        const bool terminal = (wave_.flags[k] & BMCTS_F_TERMINAL) != 0;
        // beyond the node cap the child is evaluated but neither linked nor kept: one scratch
        // node is reused, so the memory of the tree is bounded by MaxNumNodes
        leaf = link ? NewNode(parent, nullptr, wave_.next_alive[k], parent->state_depth + 1, terminal)
                    : ScratchNode(parent, wave_.next_alive[k], parent->state_depth + 1, terminal);
        for (int a = 0; a < A_; ++a)
          for (int f = 0; f < 4; ++f) leaf->pose[(size_t)a * 4 + f] = wave_.next_pose[wave_.at(a, k) * 4 + f];
        leaf->edge_reward_ego = wave_.reward[wave_.at(0, k)];
        for (size_t a = 0; a < leaf->edge_reward.size(); ++a) leaf->edge_reward[a] = wave_.reward[wave_.at((int)a, k)];
        leaf->edge_cost[0] = wave_.cost[k];
        leaf->edge_cost[1] = wave_.cost[(size_t)wave_.n + k];
        if (link) *place = leaf;
        SetHeuristic(leaf, r.ret_ego, cost, wave_.ret_agents.data(), k);
      }
    } else {
      // existing terminal node: RandomHeuristic returns a zero estimate
      const double zero[2] = {0.0, 0.0};
      SetHeuristicZero(leaf, zero);
    }
    // back-propagation child -> parent up to the root
    Node* child = leaf;
    for (int i = (int)sel.path.size() - 1; i >= 0; --i) {
      PathStep& st = sel.path[i];
      Node* n = st.node;
      if (ego_mode_ == EgoMode::kUct) {
        n->ego_uct->ReleasePending(st.ego_action);
        n->ego_uct->UpdateStatistic(st.ego_action, child->edge_reward_ego, child->LatestReturn());
      } else {
        n->ego_cc->ReleasePending(st.ego_action);
        const double child_cost[2] = {child->LatestCost(0), child->LatestCost(1)};
        n->ego_cc->UpdateStatistic(st.ego_action, child->edge_reward_ego, child->edge_cost,
                                   child->LatestReturn(), child_cost);
      }
      if (defer >= 0 && i == 0) {  // the root step (path[0].node is the root)
        defer_cost_[defer] = child->edge_cost[0];
        defer_latest_[defer] = child->latest_cost_other;
      }
      for (int s = 1; s < A_; ++s) {
        if (st.other_idx[s] < 0) continue;
        if (other_mode_ == OtherMode::kHypothesis) {
          if (defer >= 0 && i == 0) break;
          const int h = sel.hyp[s];
          n->other_hyp[s].BackupEntry(h, st.other_idx[s], child->edge_cost[0],
                                      child->latest_cost_other);
        } else {
          n->other_uct[s].ReleasePending(st.other_idx[s]);
          n->other_uct[s].UpdateStatistic(st.other_idx[s], child->edge_reward[s],
                                          child->latest_ret_other[s]);
          n->latest_ret_other[s] = n->other_uct[s].latest_return();
        }
      }
      if (other_mode_ == OtherMode::kHypothesis)
        n->latest_cost_other = child->edge_cost[0] + p_.DISCOUNT_FACTOR * child->latest_cost_other;
      child = n;
    }
  }

  // NodeStatistic::update_from_heuristic of the ego statistic (plain fields while the node has none)
  void SetEgoHeuristic(Node* n, double ret, const double* cost) {
    if (n->ego_uct) {
      n->ego_uct->UpdateFromHeuristic(ret);
    } else if (n->ego_cc) {
      n->ego_cc->UpdateFromHeuristic(ret, cost);
    } else {
      n->h_return = ret;
      n->h_cost[0] = cost[0];
      n->h_cost[1] = cost[1];
      n->h_visits++;
    }
  }

  void SetHeuristicZero(Node* n, const double* zero) {
    SetEgoHeuristic(n, 0.0, zero);
    n->latest_cost_other = 0.0;
    for (double& v : n->latest_ret_other) v = 0.0;
  }

  // CostConstrainedStatistic::update_statistic_parameters: sub-gradient step on lambda
  void UpdateLambdas() {
    const auto& c = p_.cost_constrained_statistic;
    if (root_->ego_cc->total_visits() == 0) return;
    const int greedy = root_->ego_cc->GreedyAction(0.0, c.ACTION_FILTER_FACTOR, lambdas_, gen_);
    const double range = std::max(1e-9, c.COST_UPPER_BOUND - c.COST_LOWER_BOUND);
    const double lam_max = 1.0 / (std::max(1e-6, c.TAU_GRADIENT_CLIP) * std::max(1e-6, 1.0 - p_.DISCOUNT_FACTOR));
    const double step = c.GRADIENT_UPDATE_STEP / (1.0 + 0.01 * (double)iterations_);
    for (int i = 0; i < n_costs_ && i < (int)lambdas_.size(); ++i) {
      const double q = root_->ego_cc->cost_statistic(i).pairs()[greedy].value;
      const double limit = i < (int)constraints_.size() ? constraints_[i] : 0.0;
      lambdas_[i] = std::min(lam_max, std::max(0.0, lambdas_[i] + step * (q - limit) / range));
    }
  }

  template <class F>
  void ForEachNode(const F& f) const {
    for (size_t i = 0; i < n_nodes_; ++i) f(node_chunks_[i / kChunk][i % kChunk]);
  }

  MctsParameters p_, cc_params_;
  EgoMode ego_mode_;
  OtherMode other_mode_;
  int A_, n_actions_, H_, n_costs_;
  LeafEvaluator* eval_;
  unsigned tree_index_;
  std::mt19937 gen_;
  std::vector<double> lambdas_, constraints_;
  static constexpr size_t kChunk = 256;
  std::vector<std::unique_ptr<Node[]>> node_chunks_;
  std::vector<std::unique_ptr<float[]>> pose_chunks_;
  size_t n_nodes_ = 0;
  std::unique_ptr<Node> scratch_;
  std::unique_ptr<float[]> scratch_pose_;
  Node* root_ = nullptr;
  WaveBuffers wave_;
  std::vector<Selection> sels_;
  std::vector<int> slot_;
  std::vector<double> defer_cost_, defer_latest_;  // root-level back-ups deferred within a wave
  // root-level choices of the wave in flight, [agent slot][iteration] (SelectWave / EndWave)
  std::vector<uint8_t> root_hyp_, root_new_, root_expand_;
  std::vector<int> root_idx_;
  int root_w_ = 0;
  HypothesisSampler sampler_;
  std::chrono::steady_clock::time_point t_start_;
  long wave_size_ = 1;
  int wave_in_flight_ = 0;
  bool wave_engine_ = false;
  long iterations_ = 0;
  int max_tree_depth_ = 0;
  double search_time_ms_ = 0.0;
  double t_select_ = 0.0, t_fill_ = 0.0, t_engine_ = 0.0, t_backup_ = 0.0;
  uint64_t next_draw_id_ = 1, rollout_counter_ = 0;
  NodePriorProvider* prior_ = nullptr;
  NearestObserver observer_;
  double prior_visits_ = 1.0;
  std::vector<Node*> awaiting_prior_;
  std::vector<float> obs_, pr_policy_, pr_ret_, pr_env_, pr_col_;
};

}  // namespace bark_mcts_b200
