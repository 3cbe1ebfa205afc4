#!/usr/bin/env python
"""The risk-constrained workflow of the reference end to end: a behaviour space over the IDM
parameters -> hypothesis set and scenario-risk function -> BehaviorUCTRiskConstraint, optionally
warm-started by a (here: untrained) value network evaluated on the GPU once per search wave.

    python examples/risk_constrained_rsbg.py                  # needs a B200 (there is no CPU fallback)
    python examples/risk_constrained_rsbg.py --nheuristic     # + BehaviorUCTNHeuristicRiskConstraint
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from planner_mcts_b200 import planner as P  # noqa: E402
from planner_mcts_b200 import scenarios  # noqa: E402
from planner_mcts_b200.behavior_space import BehaviorSpace  # noqa: E402
from planner_mcts_b200.nheuristic import (BehaviorUCTNHeuristicRiskConstraint,  # noqa: E402
                                          NNToValueConverterSequential)
from planner_mcts_b200.parameters import ParameterServer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nheuristic", action="store_true")
    ap.add_argument("--iterations", type=int, default=4000)
    ap.add_argument("--split", type=int, default=4, help="hypotheses per ranged parameter")
    args = ap.parse_args()

    params = ParameterServer()
    # 1. the behaviour space: headway and desired velocity are unknown, the rest is fixed
    b = params["BehaviorSpace"]["Definition"]["SpaceBoundaries"]["BehaviorIDMStochastic"]
    b["HeadwayDistribution"] = [0.5, 3.5]
    b["DesiredVelDistribution"] = [8.0, 16.0]
    for name, value in [("SpacingDistribution", 2.0), ("MaxAccDistribution", 1.7),
                        ("ComftBrakingDistribution", 1.67), ("CoolnessFactorDistribution", 0.0),
                        ("YieldingDurationDistribution", 1.0), ("NoYieldingDurationDistribution", 1.0)]:
        b[name] = [value]
    prior = params["BehaviorSpace"]["Definition"]["PriorKnowledgeFunction"]
    prior["FunctionDefinition"] = "UniformKnowledgeFunctionDefinition"
    prior["PriorKnowledgeFunction"]["NumPartitionsIntegration"] = 50
    params["BehaviorSpace"]["Hypothesis"]["BehaviorHypothesisIDM"]["NumSamples"] = 2000
    space = BehaviorSpace(params)
    hypotheses, _ = space.create_hypothesis_set_fixed_split(args.split)   # split ** 2 hypotheses
    risk_function = space.get_scenario_risk_function()
    print(f"{len(hypotheses)} hypotheses; risk normalisation constant "
          f"{risk_function.normalization_constant:.4g}")

    # 2. planner parameters (the reference's keys)
    base = params["BehaviorUctBase"]
    base["Mcts"]["MaxNumIterations"] = args.iterations
    base["Mcts"]["MaxSearchTime"] = 5000
    base["Mcts"]["State"]["SplitSafeDistCollision"] = True
    base["Mcts"]["State"]["EvaluationParameters"]["AddSafeDist"] = True
    base["EgoBehavior"]["AccelerationInputs"] = [0.0, 1.0, 4.0, -1.0, -8.0]
    base["EgoBehavior"]["AddLaneChangeActions"] = False     # a fixed action set for the value net
    params["BehaviorUctRiskConstraint"]["DefaultAvailableRisk"] = [0.1, 0.0]
    params["BehaviorUctRiskConstraint"]["EstimateScenarioRisk"] = True

    # 3. a world on the synthetic merge
    cfg, scn, meta = scenarios.workload("c3_merge_risk")
    pose, _, _ = scenarios.leaves_for(scn, meta, 1, seed=3)
    agents = [P.Agent(id=1 + a, x=float(pose[a, 0, 0]), y=float(pose[a, 0, 1]),
                      theta=float(pose[a, 0, 2]), v=float(pose[a, 0, 3])) for a in range(meta["A"])]
    world = P.ObservedWorld(ego_id=1, agents=agents, map=scn)
    base["MaxNearestAgents"] = meta["A"] - 1

    if args.nheuristic:
        import torch
        n_actions = 5
        model = torch.nn.Sequential(torch.nn.Linear(20, 64), torch.nn.ReLU(),
                                    torch.nn.Linear(64, 3 * n_actions))
        planner = BehaviorUCTNHeuristicRiskConstraint(
            params, hypotheses, risk_function, model, None,
            NNToValueConverterSequential(n_actions))
    else:
        planner = P.BehaviorUCTRiskConstraint(params, hypotheses, risk_function)

    for step in range(2):
        planner.Plan(0.2, world)
        action, policy = planner.last_policy_sampled
        print(f"step {step}: action {action} policy {policy} expected risk "
              f"{[round(r, 4) for r in planner.last_expected_risk]} available "
              f"{[round(r, 4) for r in planner.last_scenario_risk]}  "
              f"{planner.last_num_iterations} iterations in {planner.last_solution_time:.1f} ms")
        if args.nheuristic:
            calls, nodes = planner.node_prior_stats
            print(f"         model evaluated {calls} times on {nodes} nodes")
        world.world_time += 0.2


if __name__ == "__main__":
    main()
