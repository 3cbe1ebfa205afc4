#!/usr/bin/env python
"""Minimal use of the drop-in planner: an SBG planner (8 IDM hypotheses over the headway range)
on the synthetic merge scenario, one Plan() per simulation step.

    python examples/plan_merge.py                 # needs a B200 (there is no CPU fallback)
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from planner_mcts_b200 import planner as P  # noqa: E402
from planner_mcts_b200 import scenarios  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--iterations", type=int, default=2000)
    args = ap.parse_args()

    cfg, scn, meta = scenarios.workload("c2_merge_sbg")          # map + hypothesis set
    pose, alive, _ = scenarios.leaves_for(scn, meta, 1, seed=3)   # one concrete world
    agents = [P.Agent(id=1 + a, x=float(pose[a, 0, 0]), y=float(pose[a, 0, 1]),
                      theta=float(pose[a, 0, 2]), v=float(pose[a, 0, 3])) for a in range(meta["A"])]
    world = P.ObservedWorld(ego_id=1, agents=agents, map=scn)

    b = "BehaviorUctBase::"
    params = {b + "Mcts::MaxNumIterations": args.iterations, b + "Mcts::MaxSearchTime": 2000,
              b + "MaxNearestAgents": meta["A"] - 1,
              b + "EgoBehavior::AccelerationInputs": [0.0, 1.0, 4.0, -1.0, -8.0]}
    hypotheses = [P.HypothesisIDM(scn.hypotheses[h], num_samples=2000) for h in range(meta["H"])]
    planner = P.BehaviorUCTHypothesis(params, hypotheses)

    dt = 0.2
    for step in range(args.steps):
        traj = planner.Plan(dt, world)
        acc, steer = planner.last_action
        print(f"t={world.world_time:.1f}s  macro action {planner.last_macro_action}  acc {acc:+.1f} m/s2  "
              f"{planner.last_num_iterations} iterations in {planner.last_solution_time:.1f} ms  "
              f"beliefs(agent 2) {planner.current_beliefs(2).round(2).tolist()}")
        # the simulator's part: the ego follows the planned trajectory, the others keep speed
        _, x, y, th, v = traj[-1]
        ego = world.agents[0]
        ego.x, ego.y, ego.theta, ego.v, ego.last_acc = x, y, th, v, acc
        for a in world.agents[1:]:
            a.x += a.v * dt
        world.world_time += dt


if __name__ == "__main__":
    main()
