"""A small SEQUENTIAL multi-agent MCTS (TEST INFRASTRUCTURE), written from SURVEY.md Appendix B:
mamcts Mcts::search / iterate with a UctStatistic for the ego (UCB1 on normalised values, running
bound estimation, progressive widening K N^alpha) and a HypothesisStatistic per other agent
(per-hypothesis action tables, progressive widening, cost-based UCB or count-proportional
re-selection), one RandomHeuristic rollout per iteration, discounted back-propagation.

It is the plain textbook loop: dictionaries, no waves, no virtual visits, no lazily built
statistics, no arenas -- everything planner-mcts_b200/host/mcts.hpp adds for speed is absent.  With
Mcts::Engine::WaveSize = 1 the host search must visit the same actions the same number of times
(tests/test_independent_spec.py).  Leaf evaluation (one expansion step + one rollout) is a
callback; the test passes the CPU oracle.

Random numbers: std::mt19937 (numpy's legacy MT19937 seeding is the same init_genrand) with the
search's own mappings -- multiply-shift for an index, 53 bits from two draws for a uniform
double -- so that the sequence of choices is defined by the seed and nothing else."""
import math

import numpy as np


class Mt19937:
    def __init__(self, seed):
        self.rs = np.random.RandomState(int(seed) & 0xFFFFFFFF)   # init_genrand(seed)

    def next(self):
        return int(self.rs.randint(0, 1 << 32, dtype=np.uint64))

    def index(self, n):
        return (self.next() * n) >> 32

    def uniform01(self):
        hi, lo = self.next() >> 5, self.next() >> 6
        return (hi * 67108864.0 + lo) / 9007199254740992.0


class Uct:
    """ego statistic"""

    def __init__(self, n_actions, p, seed_value=0.0, seed_visits=0):
        self.count = [0] * n_actions
        self.value = [0.0] * n_actions
        self.unexpanded = list(range(n_actions))
        self.expanded = []
        self.p = p
        self.total = seed_visits           # heuristic visits the node collected before it had tables
        self.latest = seed_value

    def choose(self, gen):
        p = self.p
        n_exp = len(self.expanded)
        widen = bool(self.unexpanded) and (p["pw_alpha"] <= 0.0 or n_exp < 1 or
                                           n_exp <= p["pw_k"] * self.total ** p["pw_alpha"])
        if widen:
            a = self.unexpanded.pop(gen.index(len(self.unexpanded)))
            self.expanded.append(a)
            return a
        lo, hi = p["lower"], p["upper"]
        seen = [self.value[a] for a in range(len(self.count)) if self.count[a] > 0]
        if p["bound_estimation"] and len(seen) >= 2 and max(seen) > min(seen):
            lo, hi = min(seen), max(seen)
        two_log = 2.0 * math.log(max(1.0, float(self.total)))
        best, best_v = 0, -math.inf
        for a in range(len(self.count)):
            if a not in self.expanded:
                continue
            n = self.count[a]
            if n <= 0:
                v = math.inf
            else:
                q = (self.value[a] - lo) / (hi - lo) if hi > lo else 0.0
                v = q + 2.0 * p["c"] * math.sqrt(two_log / n)
            if v > best_v:
                best, best_v = a, v
        return best

    def update(self, a, reward, child_latest):
        self.latest = reward + self.p["gamma"] * child_latest
        self.count[a] += 1
        self.value[a] += (self.latest - self.value[a]) / self.count[a]
        self.total += 1


class HypStat:
    """another agent's statistic: per hypothesis a widening table of sampled actions"""

    def __init__(self, n_hyp, p):
        self.p = p
        self.tables = [[] for _ in range(max(1, n_hyp))]   # entries: dict(key, draw, count, cost, dead)
        self.visits = [0] * max(1, n_hyp)

    def choose(self, h, gen, fresh_draw):
        p, tab = self.p, self.tables[h]
        if p["hyp_based"]:
            n_vis, n_act = self.visits[h], len(tab)
        else:
            n_vis, n_act = sum(self.visits), sum(len(t) for t in self.tables)
        limit = (n_act / p["k"]) ** (1.0 / p["alpha"]) if p["k"] > 0 and p["alpha"] > 0 else math.inf
        idx = -1
        if not n_vis >= limit:
            if p["cost_based"]:
                explore = 2.0 * p["c"] * math.sqrt(2.0 * math.log(max(1.0, float(n_vis))))
                inv_range = 1.0 / (p["hi"] - p["lo"]) if p["hi"] > p["lo"] else 0.0
                best = -math.inf
                for i, e in enumerate(tab):
                    if e["dead"]:
                        continue
                    inv = float(np.float32(1.0) / np.sqrt(np.float32(e["count"]))) if e["count"] else 1.0e30
                    v = (e["cost"] - p["lo"]) * inv_range + explore * inv
                    if v > best:
                        best, idx = v, i
            else:
                tot = float(self.visits[h])
                if tot > 0.0:
                    u = gen.uniform01() * tot
                    for i, e in enumerate(tab):
                        if e["dead"] or e["count"] == 0:
                            continue
                        idx = i
                        u -= e["count"]
                        if u < 0.0:
                            break
        if idx < 0:
            tab.append({"key": None, "draw": fresh_draw, "count": 0, "cost": 0.0, "dead": False})
            idx = len(tab) - 1
        return idx

    def resolve(self, h, idx, key):
        tab = self.tables[h]
        if tab[idx]["key"] is not None:
            return idx
        for i, e in enumerate(tab):
            if i != idx and not e["dead"] and e["key"] == key:
                tab[idx]["dead"], tab[idx]["key"] = True, key
                return i
        tab[idx]["key"] = key
        return idx

    def update(self, h, idx, cost, child_latest_cost):
        e = self.tables[h][idx]
        latest = cost + self.p["gamma"] * child_latest_cost
        e["count"] += 1
        e["cost"] += (latest - e["cost"]) / e["count"]
        self.visits[h] += 1


class Node:
    def __init__(self, pose, alive, depth, tree_depth, terminal):
        self.pose, self.alive, self.depth, self.tree_depth, self.terminal = pose, alive, depth, tree_depth, terminal
        self.ego = None
        self.others = None
        self.children = {}
        self.h_return, self.h_cost, self.h_visits = 0.0, 0.0, 0
        self.latest_cost_other = 0.0
        self.edge_reward, self.edge_cost = 0.0, 0.0

    def latest_return(self):
        return self.ego.latest if self.ego else self.h_return


def search(root_pose, root_alive, n_agents, n_actions, n_hyp, params, iterations, sample_hypotheses,
           evaluate):
    """evaluate(pose, alive, depth, hyp, ego_action, draw_ids, rollout_id) ->
         dict(next_pose, next_alive, reward, cost0, flags_terminal, last_action[A], ret, rcost0)
    sample_hypotheses() -> list of one hypothesis id per agent slot.
    Returns the root node."""
    gen = Mt19937(params["seed"])
    up, hp = params["uct"], params["hyp"]
    root = Node(root_pose, root_alive, 1, 0, False)
    next_draw, gamma, n_evaluated = 1, up["gamma"], 0

    def ensure(n):
        if n.ego is None:
            n.ego = Uct(n_actions, up, n.h_return, n.h_visits)
            n.others = [None] + [HypStat(n_hyp, hp) for _ in range(1, n_agents)]

    ensure(root)
    for it in range(iterations):
        hyp = sample_hypotheses()
        node, path, leaf = root, [], None
        while True:
            if node.terminal:
                leaf = node
                if node.ego:
                    node.ego.latest, node.ego.total = 0.0, node.ego.total + 1
                else:
                    node.h_return, node.h_cost, node.h_visits = 0.0, 0.0, node.h_visits + 1
                node.latest_cost_other = 0.0
                break
            ensure(node)
            a = node.ego.choose(gen)
            idx, new = {}, {}
            for s in range(1, n_agents):
                if not (node.alive >> s) & 1:
                    continue
                idx[s] = node.others[s].choose(hyp[s], gen, next_draw)
                next_draw += 1
                new[s] = node.others[s].tables[hyp[s]][idx[s]]["key"] is None
            step = {"node": node, "a": a, "idx": idx}
            path.append(step)
            if not any(new.values()):
                key = (a,) + tuple(node.others[s].tables[hyp[s]][idx[s]]["key"] for s in sorted(idx))
                if key in node.children:
                    node = node.children[key]
                    continue
            draws = {s: node.others[s].tables[hyp[s]][idx[s]]["draw"] for s in idx}
            out = evaluate(node.pose, node.alive, node.depth, hyp, a, draws, n_evaluated)
            n_evaluated += 1   # rollout ids count the leaf evaluations
            for s in idx:
                if new[s]:
                    idx[s] = node.others[s].resolve(hyp[s], idx[s], out["last_action"][s])
            key = (a,) + tuple(node.others[s].tables[hyp[s]][idx[s]]["key"] for s in sorted(idx))
            child = node.children.get(key)
            if child is None:
                child = Node(out["next_pose"], out["next_alive"], node.depth + 1, node.tree_depth + 1,
                             out["terminal"])
                child.edge_reward, child.edge_cost = out["reward"], out["cost0"]
                node.children[key] = child
                child.h_return, child.h_cost, child.h_visits = out["ret"], out["rcost0"], 1
                child.latest_cost_other = out["rcost0"]
            elif child.ego is not None:
                child.ego.latest = out["ret"]
                child.latest_cost_other = out["rcost0"]
            else:
                child.h_return, child.h_cost, child.h_visits = out["ret"], out["rcost0"], child.h_visits + 1
                child.latest_cost_other = out["rcost0"]
            leaf = child
            break
        child = leaf
        for step in reversed(path):
            n = step["node"]
            n.ego.update(step["a"], child.edge_reward, child.latest_return())
            for s, i in step["idx"].items():
                n.others[s].update(hyp[s], i, child.edge_cost, child.latest_cost_other)
            n.latest_cost_other = child.edge_cost + gamma * child.latest_cost_other
            child = n
    return root
