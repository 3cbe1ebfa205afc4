"""Bias of the wave-batched search against the sequential search by wave size (CPU oracle
backend; test infrastructure, like everything that imports tests/oracle.py).

    python tools/wave_bias.py [seeds]      # profiles/r2_wave_bias_64seeds.txt was made with 64

Three planners x two worlds x 1500 / 2000 / 4000 iterations; per case the waves iterations / 128
(the default), a candidate of up to iterations / 64, and twice that.  Per line: the extreme
z-scores (pooled standard errors of the seed means) and the extreme bias in seed-to-seed standard
deviations over the root estimates, and the set of decisions taken."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle, planner_worlds as W
from test_planner import cpu, _root_estimates, P, P0
NS = int(sys.argv[1]) if len(sys.argv) > 1 else 64
def run(kind, world_kw, iters, wave, seed):
    world = W.observed_world(world_kw["n_others"], world_kw["rel"], world_kw["v"], world_kw["dv"], goal_center=(150.0, -1.75))
    params = W.base_params(iters, seed=1000 + seed)
    params[P0 + "Mcts::Engine::WaveSize"] = wave
    if kind == "hyp":
        pl = cpu(P.BehaviorUCTHypothesis, params, W.hypotheses())
    elif kind == "risk":
        params[P0 + "Mcts::RandomHeuristic::MaxNumIterations"] = 20
        params.update({P0 + "Mcts::State::SplitSafeDistCollision": 1, P0 + "Mcts::State::EvaluationParameters::AddSafeDist": 1,
                       P0 + "Mcts::State::CollisionCost": 1.0, P0 + "Mcts::State::GoalCost": 0.0,
                       "BehaviorUctRiskConstraint::DefaultAvailableRisk": [0.1, 0.0]})
        pl = cpu(P.BehaviorUCTRiskConstraint, params, W.hypotheses())
    else:
        params[P0 + "Mcts::RandomHeuristic::MaxNumIterations"] = 20
        params[P0 + "Mcts::UctStatistic::ProgressiveWidening::K"] = 4.0
        pl = cpu(P.BehaviorUCTCooperative, params)
    pl.Plan(0.2, world)
    return _root_estimates(pl, 5), pl.last_macro_action
def default_wave(iters):
    return max(1, int(iters / 128.0 * min(2.0, max(1.0, iters / 1000.0))))
W1 = dict(n_others=1, rel=3.0, v=6.0, dv=3.0); W2 = dict(n_others=2, rel=8.0, v=8.0, dv=2.0)
cases = [(k, w, it) for it in (1500, 2000, 4000) for k in ("hyp", "risk", "coop") for w in (W1, W2)]
for kind, wk, iters in cases:
    seq = np.array([run(kind, wk, iters, 1, s)[0] for s in range(NS)])
    for wave in (max(1, iters // 128), default_wave(iters), 2 * default_wave(iters)):
        r = [run(kind, wk, iters, wave, s) for s in range(NS)]
        bat = np.array([x[0] for x in r])
        ok = ~(np.isnan(seq).any(0) | np.isnan(bat).any(0))
        d = (bat.mean(0) - seq.mean(0))[ok]
        sem = np.sqrt(seq.var(0, ddof=1) / NS + bat.var(0, ddof=1) / NS)[ok]
        sd = np.sqrt(0.5 * (seq.var(0, ddof=1) + bat.var(0, ddof=1)))[ok]
        z = d / np.maximum(sem, 1e-12)
        print(f"{kind} n_others={wk['n_others']} iters={iters} wave={wave}: z min {z.min():+.2f} max {z.max():+.2f}; bias/sd min {np.min(d/np.maximum(sd,1e-12)):+.3f} max {np.max(d/np.maximum(sd,1e-12)):+.3f}; decisions {sorted(set(x[1] for x in r))}", flush=True)
