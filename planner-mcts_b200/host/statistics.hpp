// statistics.hpp -- node statistics and belief tracker of the host tree search.
//
// Restatement of mamcts (github.com/juloberno/mamcts @b3e65e7, not vendored in the
// reference; behaviour summarised in SURVEY.md Appendix B): UctStatistic,
// HypothesisStatistic, CostConstrainedStatistic, HypothesisBeliefTracker.  Every
// statistic additionally carries "pending" visit counts so that a wave of selections
// can be made before their rollouts return (virtual visits); with a wave size of 1
// they reduce to the sequential mamcts updates.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <limits>
#include <map>
#include <random>
#include <vector>

#include "params.hpp"
#include "world.hpp"

namespace bark_mcts_b200 {

using ActionKey = uint32_t;
constexpr ActionKey kNoKey = 0xFFFFFFFFu;  // a NaN pattern: never the key of an acceleration

// uniform double in [0, 1) from two draws; std::generate_canonical evaluates two long-double
// logarithms per call, which made it the most expensive part of hypothesis sampling
inline double Uniform01(std::mt19937& gen) {
  const uint64_t hi = gen() >> 5, lo = gen() >> 6;  // 27 + 26 bits
  return ((double)hi * 67108864.0 + (double)lo) * (1.0 / 9007199254740992.0);
}

// uniform index in [0, n) by multiply-shift on one 32-bit draw.  (std::uniform_int_distribution is
// implementation defined: the same seed would give different searches under another standard
// library; this mapping is part of the search's definition and tests/tiny_mcts.py mirrors it.)
inline size_t UniformIndex(std::mt19937& gen, size_t n) {
  return (size_t)(((uint64_t)gen() * (uint64_t)n) >> 32);
}

// sqrt(2 ln N) of the UCB exploration term for small N (one table for the process)
inline double SqrtTwoLog(double n) {
  constexpr int kN = 4096;
  static const std::vector<double> table = [] {
    std::vector<double> t(kN);
    for (int i = 0; i < kN; ++i) t[i] = std::sqrt(2.0 * std::log(std::max(1.0, (double)i)));
    return t;
  }();
  const long k = (long)n;
  return (n >= 0.0 && k < kN && (double)k == n) ? table[k] : std::sqrt(2.0 * std::log(std::max(1.0, n)));
}

inline ActionKey KeyOfAcceleration(float acc) {
  // the reference addresses a continuous action by boost::hash of its value
  // (action_store/behavior_action_store.hpp:62-76); the bit pattern serves the same purpose
  uint32_t b;
  std::memcpy(&b, &acc, 4);
  return (ActionKey)b;
}

// ------------------------------------------------------------------ UctStatistic
class UctStatistic {
 public:
  struct Pair {
    unsigned count = 0;
    unsigned pending = 0;
    double value = 0.0;
    bool expanded = false;
  };

  // pw_k / pw_alpha: Mcts::UctStatistic::ProgressiveWidening::{K, Alpha}
  // (param_config/mcts_parameters_from_param_server.hpp:63-64); pw_alpha <= 0 switches the
  // widening off (every action is expanded before UCB is used)
  void Init(int num_actions, double lower, double upper, double exploration, double gamma,
            bool bound_estimation, double pw_k = 1.0, double pw_alpha = 0.0) {
    pw_k_ = pw_k;
    pw_alpha_ = pw_alpha;
    pairs_.assign(num_actions, Pair());
    lower_ = lower;
    upper_ = upper;
    c_ = exploration;
    gamma_ = gamma;
    bound_estimation_ = bound_estimation;
    unexpanded_.resize(num_actions);
    for (int i = 0; i < num_actions; ++i) unexpanded_[i] = i;
  }

  int num_actions() const { return (int)pairs_.size(); }
  bool all_expanded() const { return unexpanded_.empty(); }
  unsigned total_visits() const { return total_visits_; }
  double value() const { return value_; }
  double latest_return() const { return latest_return_; }
  const std::vector<Pair>& pairs() const { return pairs_; }

  // progressive widening of the ego / cooperative action set: a new action is opened only while
  // #expanded <= K * N^alpha (the form HypothesisStatistic uses per hypothesis); with the
  // reference defaults K = 1, alpha = 0.1 that is two actions until 1024 visits
  bool WideningAllowed() const {
    if (unexpanded_.empty()) return false;
    if (!(pw_alpha_ > 0.0)) return true;
    const double n_exp = (double)(pairs_.size() - unexpanded_.size());
    if (n_exp < 1.0) return true;
    return n_exp <= pw_k_ * std::pow((double)(total_visits_ + total_pending_), pw_alpha_);
  }

  // mamcts UctStatistic::choose_next_action: a random unexpanded action while progressive
  // widening allows one, else UCB1 over the expanded actions
  int ChooseNextAction(std::mt19937& gen) {
    int a;
    if (WideningAllowed()) {
      size_t i = UniformIndex(gen, unexpanded_.size());
      a = unexpanded_[i];
      unexpanded_.erase(unexpanded_.begin() + i);
      pairs_[a].expanded = true;
    } else {
      a = UcbMaximizingAction();
    }
    pairs_[a].pending++;
    total_pending_++;
    return a;
  }

  int UcbMaximizingAction() const {
    double lo = lower_, hi = upper_;
    EstimatedBounds(&lo, &hi);
    const double n_total = std::max(1.0, (double)(total_visits_ + total_pending_));
    const double two_log_n = 2.0 * std::log(n_total);
    int best = 0;
    double best_v = -std::numeric_limits<double>::infinity();
    for (int a = 0; a < (int)pairs_.size(); ++a) {
      const Pair& p = pairs_[a];
      if (!p.expanded) continue;
      const double n = (double)(p.count + p.pending);
      double v;
      if (n <= 0.0) {
        v = std::numeric_limits<double>::max();
      } else if (p.count == 0) {
        // only visits in flight (same wave): nothing is known about the action yet, and its
        // initial value of 0 must not attract the rest of the wave.  Chosen again only if every
        // expanded action is in this state.
        v = -std::numeric_limits<double>::max();
      } else {
        const double q = (hi > lo) ? (p.value - lo) / (hi - lo) : 0.0;
        v = q + 2.0 * c_ * std::sqrt(two_log_n / n);
      }
      if (v > best_v) {
        best_v = v;
        best = a;
      }
    }
    return best;
  }

  void ReleasePending(int a) {
    if (pairs_[a].pending) pairs_[a].pending--;
    if (total_pending_) total_pending_--;
  }

  // what `visits` heuristic updates, the last one with this value, have left behind (a node
  // keeps its heuristic estimates in plain fields until it needs its action tables)
  void SeedHeuristic(double heuristic_value, unsigned visits) {
    value_ = heuristic_value;
    latest_return_ = heuristic_value;
    total_visits_ = visits;
  }

  // mamcts update_from_heuristic
  void UpdateFromHeuristic(double heuristic_value) {
    value_ = heuristic_value;
    latest_return_ = heuristic_value;
    total_visits_ += 1;
  }

  // a fresh rollout from a node that already has action statistics (a newly planned joint
  // action resolved to this existing child): the sample is what the parents back up, the
  // node's own value estimate and visit count stay as they are
  void SetLatestReturn(double ret) { latest_return_ = ret; }

  // mamcts update_statistic(changed_child_statistic)
  void UpdateStatistic(int action, double collected_reward, double child_latest_return) {
    Pair& p = pairs_[action];
    latest_return_ = collected_reward + gamma_ * child_latest_return;
    p.count += 1;
    p.value += (latest_return_ - p.value) / p.count;
    total_visits_ += 1;
    value_ += (latest_return_ - value_) / total_visits_;
  }

  int BestAction() const {  // highest action value among visited actions
    int best = 0;
    double bv = -std::numeric_limits<double>::infinity();
    for (int a = 0; a < (int)pairs_.size(); ++a)
      if (pairs_[a].count > 0 && pairs_[a].value > bv) {
        bv = pairs_[a].value;
        best = a;
      }
    return best;
  }

  Policy GetPolicy() const {  // action -> action value ("last_return_values")
    Policy pol;
    for (int a = 0; a < (int)pairs_.size(); ++a)
      if (pairs_[a].count > 0) pol[(unsigned)a] = pairs_[a].value;
    return pol;
  }

  double Normalized(int a) const {
    return (upper_ > lower_) ? (pairs_[a].value - lower_) / (upper_ - lower_) : 0.0;
  }

  // root-parallel merge (mamcts merge_node_statistics): visit-weighted means
  void MergeRaw(int a, double count, double value_sum) {
    Pair& p = pairs_[a];
    p.count = (unsigned)count;
    p.value = count > 0 ? value_sum / count : 0.0;
    p.expanded = p.expanded || count > 0;
  }

 private:
  void EstimatedBounds(double* lo, double* hi) const {
    if (!bound_estimation_) return;
    double mn = std::numeric_limits<double>::infinity(), mx = -mn;
    int n = 0;
    for (const Pair& p : pairs_)
      if (p.count > 0) {
        mn = std::min(mn, p.value);
        mx = std::max(mx, p.value);
        ++n;
      }
    if (n >= 2 && mx > mn) {
      *lo = mn;
      *hi = mx;
    }
  }

  std::vector<Pair> pairs_;
  std::vector<int> unexpanded_;
  double lower_ = 0, upper_ = 1, c_ = 0.7, gamma_ = 0.9;
  double pw_k_ = 1.0, pw_alpha_ = 0.0;
  bool bound_estimation_ = true;
  unsigned total_visits_ = 0, total_pending_ = 0;
  double value_ = 0.0, latest_return_ = 0.0;
};

// -------------------------------------------------------------- HypothesisStatistic
// Another agent's node statistic: per hypothesis a widening set of sampled actions
// (each = one IDM parameter draw) with visit counts and the ego cost they led to.
class HypothesisStatistic {
 public:
  // One expanded action of one hypothesis.  What the selection scan reads at every visit lives
  // in Hot (16 bytes, contiguous per table); the rest is only touched for the chosen entry.
  struct Entry {
    ActionKey key = 0;      // bit pattern of the applied acceleration, once resolved
    uint64_t draw_id = 0;   // Philox id that reproduces the action (replaces the stored trajectory)
    unsigned count = 0, pending = 0;
    bool resolved = false;
    bool dead = false;      // merged into another entry with the same action: never selected
  };
  struct Hot {
    double ego_cost = 0.0;     // mean accumulated ego cost after this action
    float inv_sqrt_n = 1.0f;   // 1 / sqrt(n); -inf for a dead entry (never the UCB maximum)
    uint32_t n = 0;            // count + pending; 0 for a dead entry
  };

  // everything of one hypothesis behind one header.  (Software prefetch of the tables of all
  // agents ahead of a selection step was measured: 8 % slower than leaving it to the core.)
  struct Table {
    std::vector<Hot> hot;
    std::vector<Entry> cold;
    std::vector<ActionKey> keys;  // key of a resolved live entry, else kNoKey
    unsigned visits = 0, pending = 0;
    double widen_n = -1.0, widen_v = 0.0;  // progressive-widening threshold for widen_n actions
  };

  void Init(int num_hypotheses, const MctsParameters& m) {
    tables_.assign(std::max(1, num_hypotheses), Table());
    k_ = m.hypothesis_statistic.PROGRESSIVE_WIDENING_K;
    alpha_ = m.hypothesis_statistic.PROGRESSIVE_WIDENING_ALPHA;
    hyp_based_ = m.hypothesis_statistic.PROGRESSIVE_WIDENING_HYPOTHESIS_BASED;
    cost_based_ = m.hypothesis_statistic.COST_BASED_ACTION_SELECTION;
    lo_ = m.hypothesis_statistic.LOWER_COST_BOUND;
    hi_ = m.hypothesis_statistic.UPPER_COST_BOUND;
    c_ = m.hypothesis_statistic.EXPLORATION_CONSTANT;
    gamma_ = m.DISCOUNT_FACTOR;
  }

  // returns the index of the chosen entry in the table of hypothesis h; *is_new tells the
  // caller to plan the action (S::plan_action_current_hypothesis) with entry.draw_id
  int ChooseNextAction(int h, std::mt19937& gen, uint64_t fresh_draw_id, bool* is_new) {
    Table& t = tables_[h];
    std::vector<Entry>& tab = t.cold;
    std::vector<Hot>& hot = t.hot;
    double n_vis, n_act;
    if (hyp_based_) {
      n_vis = t.visits + t.pending;
      n_act = (double)tab.size();
    } else {
      n_vis = 0;
      n_act = 0;
      for (const Table& o : tables_) {
        n_vis += o.visits + o.pending;
        n_act += (double)o.cold.size();
      }
    }
    // progressive widening: |actions| <= K * N^alpha.  Inside a wave an entry may be chosen
    // again before its action is known (unresolved): it is re-planned with the same draw id,
    // which yields the same action and therefore the same child.
    // n_act <= K * N^alpha  <=>  N >= (n_act / K)^(1 / alpha); the right-hand side only
    // changes when an action is added, so it is cached (pow is the most expensive call of a
    // selection step)
    // (kept per table: with hypothesis-based widening the tables of one agent differ in size,
    // and a cache per statistic would be recomputed whenever the sampled hypothesis changes)
    if (n_act != t.widen_n) {
      t.widen_n = n_act;
      t.widen_v = (k_ > 0.0 && alpha_ > 0.0) ? std::pow(n_act / k_, 1.0 / alpha_)
                                             : std::numeric_limits<double>::infinity();
    }
    bool widen = n_vis >= t.widen_v;
    int idx = -1;
    if (!widen) {
      const int n_tab = (int)hot.size();
      if (cost_based_) {
        // worst case for the ego: UCB on the normalised ego cost (RMDP / RSBG),
        // v = normalised cost + 2 c sqrt(2 ln N / n)
        double best = -std::numeric_limits<double>::infinity();
        const double explore = 2.0 * c_ * SqrtTwoLog(n_vis);
        const double inv_range = hi_ > lo_ ? 1.0 / (hi_ - lo_) : 0.0;
        for (int i = 0; i < n_tab; ++i) {
          // (selects, not branches: where the running maximum moves is not predictable)
          const double v = (hot[i].ego_cost - lo_) * inv_range + explore * (double)hot[i].inv_sqrt_n;
          const bool better = v > best;
          best = better ? v : best;
          idx = better ? i : idx;
        }
      } else {
        // stochastic behaviour model: re-sample among the expanded actions by visit count.
        // Every live entry has been chosen at least once and merged entries carry nothing, so
        // the weights sum to visits + pending of this hypothesis.
        const double tot = (double)t.visits + (double)t.pending;
        if (tot > 0.0) {
          double u = Uniform01(gen) * tot;
          for (int i = 0; i < n_tab; ++i) {
            if (hot[i].n == 0) continue;
            idx = i;
            u -= (double)hot[i].n;
            if (u < 0.0) break;
          }
        }
      }
    }
    if (idx < 0) {  // widening allowed, or nothing to choose from yet
      Entry e;
      e.draw_id = fresh_draw_id;
      tab.push_back(e);
      hot.push_back(Hot());
      t.keys.push_back(kNoKey);
      idx = (int)tab.size() - 1;
    }
    *is_new = !tab[idx].resolved;
    tab[idx].pending++;
    SetN(hot[idx], hot[idx].n + 1);
    t.pending++;
    return idx;
  }

  // the planned action of a new entry is known: resolve its key; merge with an existing
  // entry that carries the same action (the reference keys the action store by value)
  int Resolve(int h, int idx, ActionKey key) {
    Table& t = tables_[h];
    std::vector<Entry>& tab = t.cold;
    std::vector<Hot>& hot = t.hot;
    if (tab[idx].resolved) return idx;  // resolved by an earlier backup of the same wave
    const std::vector<ActionKey>& keys = t.keys;  // keys of the resolved live entries
    for (int i = 0; i < (int)keys.size(); ++i)
      if (keys[i] == key && i != idx) {
        tab[i].pending += tab[idx].pending;
        SetN(hot[i], hot[i].n + tab[idx].pending);
        tab[idx].pending = 0;
        tab[idx].dead = true;  // never selected again
        tab[idx].key = key;
        hot[idx].n = 0;
        hot[idx].inv_sqrt_n = -std::numeric_limits<float>::infinity();
        hot[idx].ego_cost = 0.0;
        return i;
      }
    tab[idx].key = key;
    tab[idx].resolved = true;
    t.keys[idx] = key;
    return idx;
  }

  void ReleasePending(int h, int idx) {
    Table& t = tables_[h];
    Entry& e = t.cold[idx];
    if (e.pending) {
      e.pending--;
      if (!e.dead) SetN(t.hot[idx], t.hot[idx].n - 1);
    }
    if (t.pending) t.pending--;
  }

  void UpdateFromHeuristic(double ego_cost) { latest_ego_cost_ = ego_cost; }

  void UpdateStatistic(int h, int idx, double collected_ego_cost, double child_latest_ego_cost) {
    Table& tb = tables_[h];
    Entry& e = tb.cold[idx];
    Hot& t = tb.hot[idx];
    latest_ego_cost_ = collected_ego_cost + gamma_ * child_latest_ego_cost;
    e.count += 1;
    t.ego_cost += (latest_ego_cost_ - t.ego_cost) / e.count;
    SetN(t, t.n + 1);
    tb.visits += 1;
  }

  // back-propagation through a selected entry: ReleasePending + UpdateStatistic in one step
  // (count + pending is unchanged, so the cached square root stays)
  void BackupEntry(int h, int idx, double collected_ego_cost, double child_latest_ego_cost) {
    Table& tb = tables_[h];
    Entry& e = tb.cold[idx];
    if (!e.pending || e.dead) {  // not the common case: fall back to the two plain steps
      ReleasePending(h, idx);
      UpdateStatistic(h, idx, collected_ego_cost, child_latest_ego_cost);
      return;
    }
    e.pending--;
    if (tb.pending) tb.pending--;
    latest_ego_cost_ = collected_ego_cost + gamma_ * child_latest_ego_cost;
    e.count += 1;
    Hot& t = tb.hot[idx];
    t.ego_cost += (latest_ego_cost_ - t.ego_cost) / e.count;
    tb.visits += 1;
  }

  double latest_ego_cost() const { return latest_ego_cost_; }
  const std::vector<Entry>& table(int h) const { return tables_[h].cold; }
  // key of a resolved live entry from the compact key array (one cache line per table)
  ActionKey KeyOf(int h, int idx) const { return tables_[h].keys[idx]; }
  double ego_cost(int h, int idx) const { return tables_[h].hot[idx].ego_cost; }
  int num_hypotheses() const { return (int)tables_.size(); }
  int NumExpandedActions() const {  // distinct resolved actions over all hypotheses
    std::vector<ActionKey> keys;
    for (const Table& t : tables_)
      for (const Entry& e : t.cold)
        if (e.resolved) keys.push_back(e.key);
    std::sort(keys.begin(), keys.end());
    return (int)(std::unique(keys.begin(), keys.end()) - keys.begin());
  }

 private:
  static void SetN(Hot& t, uint32_t n) {
    t.n = n;
    // a live entry nobody has tried or is trying comes first, as in UCB1
    // (1 / sqrt(n) from a table for the counts that occur: a division and a square root per
    // selection otherwise)
    constexpr uint32_t kN = 2048;
    static const std::vector<float> table = [] {
      std::vector<float> t(kN, 1.0e30f);
      for (uint32_t i = 1; i < kN; ++i) t[i] = 1.0f / std::sqrt((float)i);
      return t;
    }();
    t.inv_sqrt_n = n < kN ? table[n] : 1.0f / std::sqrt((float)n);
  }

  std::vector<Table> tables_;
  double k_ = 4, alpha_ = 0.25, lo_ = 0, hi_ = 1, c_ = 0.7, gamma_ = 0.9;
  bool hyp_based_ = true, cost_based_ = false;
  double latest_ego_cost_ = 0.0;
};

// --------------------------------------------------------- CostConstrainedStatistic
// Ego statistic of BehaviorUCTRiskConstraint: reward UCT + one UCT per cost channel and
// a Lagrange multiplier per constraint (cost-constrained MCTS, CC-MCP style).
class CostConstrainedStatistic {
 public:
  void Init(int num_actions, int num_costs, const MctsParameters& m) {
    const auto& c = m.cost_constrained_statistic;
    reward_.Init(num_actions, c.REWARD_LOWER_BOUND, c.REWARD_UPPER_BOUND, 0.0, m.DISCOUNT_FACTOR,
                 false);
    InitLight(num_costs);  // keeps heuristic values set before the tables were needed
    for (int i = 0; i < n_costs_; ++i)
      costs_[i].Init(num_actions, c.COST_LOWER_BOUND, c.COST_UPPER_BOUND, 0.0, m.DISCOUNT_FACTOR, false);
    kappa_ = c.KAPPA;
    filter_ = c.ACTION_FILTER_FACTOR;
    constraints_ = c.COST_CONSTRAINTS;
    constraints_.resize(num_costs, 0.0);
    step_cost_sum_.assign(num_actions, std::vector<double>(num_costs, 0.0));
    step_cost_n_.assign(num_actions, 0);
    min_visits_ready_ = c.MIN_VISITS_POLICY_READY;
    unexpanded_.resize(num_actions);
    for (int i = 0; i < num_actions; ++i) unexpanded_[i] = i;
    pending_.assign(num_actions, 0);
  }

  // what a freshly expanded node needs to receive its heuristic estimate; the action tables
  // (Init) are built when the node is first selected through
  void InitLight(int num_costs) {
    n_costs_ = std::min(num_costs, 2);
  }

  int num_actions() const { return reward_.num_actions(); }
  int num_costs() const { return n_costs_; }
  const UctStatistic& reward_statistic() const { return reward_; }
  const UctStatistic& cost_statistic(int i) const { return costs_[i]; }
  UctStatistic& reward_statistic() { return reward_; }
  UctStatistic& cost_statistic(int i) { return costs_[i]; }
  unsigned total_visits() const { return reward_.total_visits(); }

  bool policy_is_ready() const {
    if (min_visits_ready_ < 0) return unexpanded_.empty();
    return (long)reward_.total_visits() >= min_visits_ready_;
  }

  int ChooseNextAction(const std::vector<double>& lambdas, std::mt19937& gen) {
    int a;
    if (!policy_is_ready() && !unexpanded_.empty()) {
      size_t i;
      if (explore_policy_.empty()) {
        i = UniformIndex(gen, unexpanded_.size());
      } else {
        // NeuralCostConstrainedStatistic::choose_next_action: sample_policy(exploration policy),
        // restricted to the actions not tried yet
        double tot = 0.0;
        for (int u : unexpanded_) tot += explore_policy_[u];
        double r = Uniform01(gen) * tot;
        i = unexpanded_.size() - 1;
        for (size_t k = 0; k < unexpanded_.size(); ++k) {
          r -= explore_policy_[unexpanded_[k]];
          if (r < 0.0) {
            i = k;
            break;
          }
        }
        if (!(tot > 0.0)) i = UniformIndex(gen, unexpanded_.size());
      }
      a = unexpanded_[i];
      unexpanded_.erase(unexpanded_.begin() + i);
    } else {
      a = GreedyAction(kappa_, filter_, lambdas, gen);
      auto it = std::find(unexpanded_.begin(), unexpanded_.end(), a);
      if (it != unexpanded_.end()) unexpanded_.erase(it);
    }
    pending_[a]++;
    total_pending_++;
    return a;
  }

  void ReleasePending(int a) {
    if (pending_[a]) pending_[a]--;
    if (total_pending_) total_pending_--;
  }

  // greedy_policy(kappa, action_filter_factor) (behavior_uct_risk_constraint.hpp:145-153): the
  // near-optimal action set under the Lagrangian value, then the stochastic policy over it that
  // the reference finds with an OR-tools LP (CC-MCP, Lee et al. 2018, eq. 8):
  //     min_{pi, xi >= 0}  sum_i lambda_i xi_i   s.t.  sum_a pi_a Q_Ci(a) <= c_i + xi_i  for EVERY
  //     cost constraint i,  sum_a pi_a = 1,  pi >= 0,
  // ties broken towards the higher Lagrangian value.  A vertex of this LP mixes at most
  // 1 + (number of tight constraints) actions, so with two constraints the optimum is among: one
  // action; two actions with one constraint tight; three actions with both tight -- enumerated
  // exactly, no solver.
  struct Greedy {
    int action = 0;
    int n = 0;  // size of the policy's support; 0: uniform over all actions
    int a[3] = {-1, -1, -1};
    double p[3] = {0.0, 0.0, 0.0};
  };
  Greedy GreedyCore(double kappa, double filter, const std::vector<double>& lambdas,
                    std::mt19937& gen) const {
    const int n = num_actions();
    const double n_total = std::max(2.0, (double)(reward_.total_visits() + total_pending_));
    const double ninf = -std::numeric_limits<double>::infinity();
    double ucb[BMCTS_MAX_EGO_ACTIONS], cnt[BMCTS_MAX_EGO_ACTIONS];
    const double log_total = kappa > 0 ? std::log(n_total) : 0.0;  // kappa = 0: pure greedy
    int a_star = -1;
    for (int a = 0; a < n; ++a) {
      ucb[a] = ninf;
      cnt[a] = reward_.pairs()[a].count + pending_[a];
      if (reward_.pairs()[a].count == 0) {
        if (kappa > 0 && cnt[a] == 0) {  // never tried: explore first
          ucb[a] = std::numeric_limits<double>::max();
          if (a_star < 0 || ucb[a] > ucb[a_star]) a_star = a;
        }
        continue;
      }
      double v = reward_.Normalized(a);
      for (int i = 0; i < num_costs(); ++i)
        v -= (i < (int)lambdas.size() ? lambdas[i] : 0.0) * costs_[i].Normalized(a);
      ucb[a] = kappa > 0 ? v + kappa * std::sqrt(log_total / std::max(1.0, cnt[a])) : v;
      if (a_star < 0 || ucb[a] > ucb[a_star]) a_star = a;
    }
    Greedy g;
    if (a_star < 0) {  // nothing visited yet: uniform
      g.action = (int)UniformIndex(gen, (size_t)n);
      return g;
    }
    auto single = [&](int a) {
      g.n = 1;
      g.a[0] = g.action = a;
      g.p[0] = 1.0;
    };
    single(a_star);
    if (reward_.pairs()[a_star].count == 0) return g;
    const int nc = n_costs_;
    double lim[2] = {0.0, 0.0}, w[2] = {0.0, 0.0};
    for (int i = 0; i < nc; ++i) {
      lim[i] = i < (int)constraints_.size() ? constraints_[i] : 0.0;
      w[i] = std::max(1e-6, i < (int)lambdas.size() ? lambdas[i] : 0.0);
    }
    auto qc = [&](int i, int a) { return costs_[i].pairs()[a].value; };
    auto violation = [&](const double* c) {
      double v = 0.0;
      for (int i = 0; i < nc; ++i) v += w[i] * std::max(0.0, c[i] - lim[i]);
      return v;
    };
    {  // the maximising action meets every constraint: nothing to mix (the common case)
      const double c[2] = {nc > 0 ? qc(0, a_star) : 0.0, nc > 1 ? qc(1, a_star) : 0.0};
      if (nc == 0 || violation(c) <= 0.0) return g;
    }
    // near-optimal actions (ACTION_FILTER_FACTOR), the maximising one first
    auto conf = [&](int a) { return std::sqrt(std::log(std::max(2.0, cnt[a])) / std::max(1.0, cnt[a])); };
    const double conf_star = conf(a_star);
    int near[BMCTS_MAX_EGO_ACTIONS], n_near = 0;
    near[n_near++] = a_star;
    for (int a = 0; a < n; ++a) {
      if (a == a_star || reward_.pairs()[a].count == 0) continue;
      if (std::fabs(ucb[a] - ucb[a_star]) <= filter * (conf(a) + conf_star)) near[n_near++] = a;
    }
    if (n_near < 2) return g;
    double best_v = std::numeric_limits<double>::infinity(), best_val = ninf;
    auto consider = [&](int k, const int* acts, const double* probs) {
      double c[2] = {0.0, 0.0}, val = 0.0;
      for (int j = 0; j < k; ++j) {
        for (int i = 0; i < nc; ++i) c[i] += probs[j] * qc(i, acts[j]);
        val += probs[j] * ucb[acts[j]];
      }
      const double v = violation(c);
      if (v < best_v - 1e-12 || (v <= best_v + 1e-12 && val > best_val + 1e-12)) {
        best_v = v;
        best_val = val;
        g.n = k;
        for (int j = 0; j < 3; ++j) {
          g.a[j] = j < k ? acts[j] : -1;
          g.p[j] = j < k ? probs[j] : 0.0;
        }
      }
    };
    for (int x = 0; x < n_near; ++x) {
      const double one = 1.0;
      consider(1, &near[x], &one);
    }
    for (int x = 0; x < n_near; ++x)
      for (int y = x + 1; y < n_near; ++y)
        for (int i = 0; i < nc; ++i) {  // two actions, constraint i tight
          const double cx = qc(i, near[x]), cy = qc(i, near[y]);
          if (cx == cy) continue;
          const double px = (lim[i] - cy) / (cx - cy);
          if (!(px > 0.0 && px < 1.0)) continue;
          const int acts[2] = {near[x], near[y]};
          const double probs[2] = {px, 1.0 - px};
          consider(2, acts, probs);
        }
    if (nc > 1)
      for (int x = 0; x < n_near; ++x)
        for (int y = x + 1; y < n_near; ++y)
          for (int z = y + 1; z < n_near; ++z) {  // three actions, both constraints tight
            const double a00 = qc(0, near[x]) - qc(0, near[z]), a01 = qc(0, near[y]) - qc(0, near[z]);
            const double a10 = qc(1, near[x]) - qc(1, near[z]), a11 = qc(1, near[y]) - qc(1, near[z]);
            const double det = a00 * a11 - a01 * a10;
            if (std::fabs(det) < 1e-300) continue;
            const double b0 = lim[0] - qc(0, near[z]), b1 = lim[1] - qc(1, near[z]);
            const double px = (b0 * a11 - a01 * b1) / det, py = (a00 * b1 - b0 * a10) / det;
            const double pz = 1.0 - px - py;
            if (!(px > 0.0 && py > 0.0 && pz > 0.0)) continue;
            const int acts[3] = {near[x], near[y], near[z]};
            const double probs[3] = {px, py, pz};
            consider(3, acts, probs);
          }
    g.action = g.a[0];
    if (g.n > 1) {  // sample_policy
      double u = Uniform01(gen);
      g.action = g.a[g.n - 1];
      for (int j = 0; j < g.n; ++j) {
        u -= g.p[j];
        if (u < 0.0) {
          g.action = g.a[j];
          break;
        }
      }
    }
    return g;
  }
  int GreedyAction(double kappa, double filter, const std::vector<double>& lambdas,
                   std::mt19937& gen) const {
    return GreedyCore(kappa, filter, lambdas, gen).action;
  }
  std::pair<int, Policy> GreedyPolicy(double kappa, double filter,
                                      const std::vector<double>& lambdas, std::mt19937& gen) const {
    const Greedy g = GreedyCore(kappa, filter, lambdas, gen);
    Policy pol;
    if (g.n == 0) {
      for (int a = 0; a < num_actions(); ++a) pol[(unsigned)a] = 1.0 / num_actions();
    } else {
      for (int j = 0; j < g.n; ++j) pol[(unsigned)g.a[j]] = g.p[j];
    }
    return {g.action, pol};
  }

  std::vector<double> ExpectedPolicyCost(const Policy& pol) const {
    std::vector<double> e(num_costs(), 0.0);
    for (const auto& kv : pol)
      for (int i = 0; i < num_costs(); ++i) e[i] += kv.second * costs_[i].pairs()[kv.first].value;
    return e;
  }

  // calc_updated_constraints_based_on_policy: the budget left after executing the sampled
  // action = its expected cumulative cost plus the unused slack, minus the mean cost of the
  // step itself, carried one discount step forward
  std::vector<double> UpdatedConstraints(const std::pair<int, Policy>& sampled,
                                         const std::vector<double>& current, double gamma) const {
    std::vector<double> expected = ExpectedPolicyCost(sampled.second);
    std::vector<double> out(current.size(), 0.0);
    const int a = sampled.first;
    for (size_t i = 0; i < current.size() && (int)i < num_costs(); ++i) {
      const double slack = std::max(0.0, current[i] - expected[i]);
      const double step = step_cost_n_[a] ? step_cost_sum_[a][i] / step_cost_n_[a] : 0.0;
      out[i] = (costs_[i].pairs()[a].value + slack - step) / std::max(gamma, 1e-6);
    }
    return out;
  }

  void UpdateFromHeuristic(double ret, const double* cost) {
    reward_.UpdateFromHeuristic(ret);
    for (int i = 0; i < num_costs(); ++i) costs_[i].UpdateFromHeuristic(cost[i]);
  }

  void SetLatest(double ret, const double* cost) {
    reward_.SetLatestReturn(ret);
    for (int i = 0; i < num_costs(); ++i) costs_[i].SetLatestReturn(cost[i]);
  }

  void SeedHeuristic(double ret, const double* cost, unsigned visits) {
    reward_.SeedHeuristic(ret, visits);
    for (int i = 0; i < num_costs(); ++i) costs_[i].SeedHeuristic(cost[i], visits);
  }
  double latest_return() const { return reward_.latest_return(); }
  double latest_cost(int i) const { return costs_[i].latest_return(); }

  // update_statistic(changed_child_statistic): the child's latest return and costs
  void UpdateStatistic(int action, double reward, const double* cost, double child_return,
                       const double* child_cost) {
    reward_.UpdateStatistic(action, reward, child_return);
    for (int i = 0; i < num_costs(); ++i) {
      costs_[i].UpdateStatistic(action, cost[i], child_cost[i]);
      step_cost_sum_[action][i] += cost[i];
    }
    step_cost_n_[action]++;
  }

  void set_constraints(const std::vector<double>& c) {
    constraints_ = c;
    constraints_.resize(num_costs(), 0.0);
  }

  // NeuralCostConstrainedStatistic::initialize_from_neural_model
  // (mcts_neural_cost_constrained_statistic.hpp:175-198), after Init: the policy becomes the
  // exploration policy; the value triple {return, envelope risk, collision risk} initialises
  // every action's means as `prior_visits` virtual visits (mamcts' set_heuristic_estimate over
  // action maps is not in the container; the weight of the prior is this engine's parameter
  // Mcts::NeuralStatistic::PriorVisits)
  void SetPrior(const float* policy, const float* ret, const float* envelope_risk,
                const float* collision_risk, double prior_visits) {
    const int n = num_actions();
    if (policy) {
      explore_policy_.assign(n, 0.0);
      for (int a = 0; a < n; ++a) explore_policy_[a] = std::max(0.0, (double)policy[a]);
    }
    if (ret && envelope_risk && collision_risk && prior_visits > 0.0) {
      for (int a = 0; a < n; ++a) {
        reward_.MergeRaw(a, prior_visits, prior_visits * (double)ret[a]);
        if (num_costs() > 0)
          costs_[0].MergeRaw(a, prior_visits, prior_visits * (double)envelope_risk[a]);
        if (num_costs() > 1)
          costs_[1].MergeRaw(a, prior_visits, prior_visits * (double)collision_risk[a]);
      }
    }
  }
  const std::vector<double>& exploration_policy() const { return explore_policy_; }

 private:
  UctStatistic reward_;
  UctStatistic costs_[2];  // envelope / collision channel (get_num_costs() is 1 or 2): inline, a
                           // node is created per iteration
  int n_costs_ = 0;
  std::vector<double> constraints_;
  std::vector<std::vector<double>> step_cost_sum_;
  std::vector<unsigned> step_cost_n_;
  std::vector<int> unexpanded_;
  std::vector<double> explore_policy_;  // empty = uniform over the unexpanded actions
  std::vector<unsigned> pending_;
  unsigned total_pending_ = 0;
  double kappa_ = 10.0, filter_ = 0.5;
  long min_visits_ready_ = -1;
};

// ----------------------------------------------------------- HypothesisBeliefTracker
class HypothesisBeliefTracker {
 public:
  explicit HypothesisBeliefTracker(const MctsParameters& m)
      : p_(m.hypothesis_belief_tracker), gen_(m.hypothesis_belief_tracker.RANDOM_SEED_HYPOTHESIS_SAMPLING) {}

  bool beliefs_initialized() const { return !beliefs_.empty(); }
  const Beliefs& get_beliefs() const { return beliefs_; }
  void update_fixed_hypothesis_set(const std::map<AgentId, unsigned>& fixed) { fixed_ = fixed; }

  // belief_update(prev, cur): probs[agent][h] = prev.get_probability(h, agent, cur.last_action)
  void belief_update(const std::map<AgentId, std::vector<double>>& probs) {
    for (const auto& kv : probs) {
      const AgentId id = kv.first;
      const size_t H = kv.second.size();
      auto& hist = history_[id];
      if (hist.size() != H) hist.assign(H, std::deque<double>());
      for (size_t h = 0; h < H; ++h) {
        hist[h].push_back(kv.second[h]);
        while (hist[h].size() > p_.HISTORY_LENGTH) hist[h].pop_front();
      }
      std::vector<double> post(H, 0.0);
      double sum = 0.0;
      for (size_t h = 0; h < H; ++h) {
        const double prior = 1.0 / (double)H;
        double v;
        if (p_.POSTERIOR_TYPE == 0) {  // product of the remembered likelihoods
          v = prior;
          for (double q : hist[h]) v *= q;
        } else {  // discounted sum, newest weighted 1
          v = 0.0;
          double w = 1.0;
          for (auto it = hist[h].rbegin(); it != hist[h].rend(); ++it) {
            v += w * (*it);
            w *= p_.PROBABILITY_DISCOUNT;
          }
          v *= prior;
        }
        post[h] = v;
        sum += v;
      }
      for (size_t h = 0; h < H; ++h) post[h] = sum > 0.0 ? post[h] / sum : 1.0 / (double)H;
      beliefs_[id] = post;
    }
  }

  // one hypothesis per agent from its belief (root-belief sampling)
  const std::map<AgentId, unsigned>& sample_current_hypothesis() {
    for (const auto& kv : beliefs_) {
      auto f = fixed_.find(kv.first);
      if (f != fixed_.end()) {
        current_[kv.first] = f->second;
        continue;
      }
      std::discrete_distribution<unsigned> d(kv.second.begin(), kv.second.end());
      current_[kv.first] = d(gen_);
    }
    return current_;
  }

  // Root-belief sampling of one tree, laid out for the selection loop: per agent slot a
  // cumulative belief table (or the fixed hypothesis), its own generator, no look-ups by id.
  class SlotSampler {
   public:
    SlotSampler(const HypothesisBeliefTracker& t, const std::vector<AgentId>& slot_ids,
                unsigned num_hypotheses, unsigned seed)
        : gen_(seed), H_(std::max(1u, num_hypotheses)) {
      slots_.resize(slot_ids.size());
      for (size_t s = 1; s < slot_ids.size(); ++s) {
        Slot& sl = slots_[s];
        auto f = t.fixed_.find(slot_ids[s]);
        if (f != t.fixed_.end()) {
          sl.fixed = (int)f->second;
          continue;
        }
        auto it = t.beliefs_.find(slot_ids[s]);
        if (it == t.beliefs_.end() || it->second.empty())
          BuildAlias(std::vector<double>(H_, 1.0), &sl);  // no belief yet: uniform over H
        else
          BuildAlias(it->second, &sl);
      }
    }
    void operator()(std::vector<uint8_t>* hyp) {
      for (size_t s = 1; s < slots_.size(); ++s) {
        const Slot& sl = slots_[s];
        if (sl.fixed >= 0) {
          (*hyp)[s] = (uint8_t)sl.fixed;
          continue;
        }
        // alias method: ONE 32-bit draw picks a column (multiply-shift) and, with the low half
        // of the product as the fraction, a side of it
        const uint64_t m = (uint64_t)gen_() * (uint64_t)sl.prob.size();
        const size_t col = (size_t)(m >> 32);
        const double frac = (double)(uint32_t)m * (1.0 / 4294967296.0);
        (*hyp)[s] = frac < sl.prob[col] ? (uint8_t)col : sl.alias[col];
      }
    }

   private:
    struct Slot {
      int fixed = -1;
      std::vector<double> prob;    // alias table over the hypotheses (Vose)
      std::vector<uint8_t> alias;
    };
    static void BuildAlias(const std::vector<double>& weights, Slot* sl) {
      const size_t n = weights.size();
      double tot = 0.0;
      for (double w : weights) tot += w > 0.0 ? w : 0.0;
      sl->prob.assign(n, 1.0);
      sl->alias.resize(n);
      std::vector<double> scaled(n);
      std::vector<size_t> small, large;
      for (size_t i = 0; i < n; ++i) {
        sl->alias[i] = (uint8_t)i;
        scaled[i] = tot > 0.0 ? std::max(0.0, weights[i]) * (double)n / tot : 1.0;
        (scaled[i] < 1.0 ? small : large).push_back(i);
      }
      while (!small.empty() && !large.empty()) {
        const size_t lo = small.back(), hi = large.back();
        small.pop_back();
        sl->prob[lo] = scaled[lo];
        sl->alias[lo] = (uint8_t)hi;
        scaled[hi] -= 1.0 - scaled[lo];
        if (scaled[hi] < 1.0) {
          large.pop_back();
          small.push_back(hi);
        }
      }
    }
    std::mt19937 gen_;
    unsigned H_;
    std::vector<Slot> slots_;
  };
  SlotSampler SamplerForSlots(const std::vector<AgentId>& slot_ids, unsigned num_hypotheses,
                              unsigned tree_index) const {
    return SlotSampler(*this, slot_ids, num_hypotheses,
                       (unsigned)p_.RANDOM_SEED_HYPOTHESIS_SAMPLING + 7919u * tree_index);
  }

 private:
  decltype(MctsParameters::hypothesis_belief_tracker) p_;
  std::mt19937 gen_;
  Beliefs beliefs_;
  std::map<AgentId, std::vector<std::deque<double>>> history_;
  std::map<AgentId, unsigned> fixed_, current_;
};

}  // namespace bark_mcts_b200
