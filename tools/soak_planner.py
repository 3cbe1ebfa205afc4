#!/usr/bin/env python
"""Soak run of the planners on the GPU: a few hundred Plan() calls over changing worlds (agents
come and go, so the engines are re-configured), planner kinds and tree counts; checks that
device memory and host RSS stay flat and that nothing errors -- run on the GPU box."""
import os
import resource
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np  # noqa: E402
import torch  # noqa: E402

from planner_mcts_b200 import planner as P  # noqa: E402
from planner_mcts_b200 import scenarios  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 300
cfg, scn, meta = scenarios.workload("c3_merge_risk")
pose, _, _ = scenarios.leaves_for(scn, meta, 1, seed=3)
b = "BehaviorUctBase::"
rng = np.random.default_rng(0)


def params(trees, threaded):
    return {b + "Mcts::MaxNumIterations": 512, b + "Mcts::MaxSearchTime": 1e9,
            b + "MaxNearestAgents": meta["A"] - 1, b + "Mcts::RandomHeuristic::MaxNumIterations": 20,
            b + "Mcts::NumParallelMcts": trees, b + "Mcts::UseMultiThreading": threaded,
            b + "Mcts::State::EvaluationParameters::AddSafeDist": 1,
            b + "Mcts::State::SplitSafeDistCollision": 1,
            "BehaviorUctRiskConstraint::DefaultAvailableRisk": [0.1, 0.0]}


hyps = [P.HypothesisIDM(scn.hypotheses[h], num_samples=500) for h in range(meta["H"])]
planners = [P.BehaviorUCTHypothesis(params(1, 0), hyps), P.BehaviorUCTRiskConstraint(params(6, 0), hyps),
            P.BehaviorUCTRiskConstraint(params(40, 1), hyps), P.BehaviorUCTCooperative(params(9, 1))]
free0 = rss0 = None
t0 = time.time()
for r in range(rounds):
    n = int(rng.integers(2, meta["A"] + 1))              # the ego and n - 1 others this round
    agents = [P.Agent(id=1 + a, x=float(pose[a, 0, 0]) + 0.3 * r % 7, y=float(pose[a, 0, 1]),
                      theta=float(pose[a, 0, 2]), v=float(pose[a, 0, 3])) for a in range(n)]
    world = P.ObservedWorld(ego_id=1, agents=agents, map=scn, world_time=0.2 * r)
    pl = planners[r % len(planners)]
    traj = pl.Plan(0.2, world)
    assert traj.shape[1] == 5 and np.isfinite(traj).all()
    assert pl.last_num_iterations > 0
    if r == 4 * len(planners):                            # after every planner has warmed up
        free0 = torch.cuda.mem_get_info()[0]
        rss0 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
free1 = torch.cuda.mem_get_info()[0]
rss1 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
print(f"{rounds} Plan() calls in {time.time() - t0:.1f} s; device memory drift "
      f"{(free0 - free1) / 2**20:.1f} MiB, host max RSS {rss0 / 1024:.0f} -> {rss1 / 1024:.0f} MiB")
assert free0 - free1 < 64 << 20, "device memory keeps growing"
assert rss1 < 1.5 * rss0 + (256 << 10), "host memory keeps growing"
print("soak ok")
