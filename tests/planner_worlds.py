"""Observed worlds of the reference's planner tests (BARK make_test_observed_world on the
one-road-two-lanes map; /root/reference/bark_mcts/models/behavior/tests/) as POD worlds for
the BehaviorUCT* planners of this repo."""
import kat_worlds
from planner_mcts_b200 import planner as P

P0 = "BehaviorUctBase::"


def base_params(iterations, acc_inputs=(0, 1, 4, -1, -8), search_time_ms=1e9, seed=1000):
    """parameters the reference's tests set (behavior_uct_hypothesis_test.cc:243-262)"""
    return {
        P0 + "Mcts::MaxNumIterations": iterations,
        P0 + "Mcts::MaxSearchTime": search_time_ms,
        P0 + "Mcts::RandomSeed": seed,
        P0 + "EgoBehavior::AccelerationInputs": list(acc_inputs),
        P0 + "EgoBehavior::AddLaneChangeActions": 0,
        P0 + "Mcts::State::GoalReward": 200.0,
        P0 + "Mcts::State::CollisionReward": -2000.0,
        P0 + "Mcts::State::GoalCost": 10.0,
        P0 + "Mcts::State::CollisionCost": 300.0,
        P0 + "MaxNearestAgents": 5,
        # :253-254 (the defaults, K = 1 and alpha = 0.1, keep the ego at two actions for the
        # first 1024 visits)
        P0 + "Mcts::UctStatistic::ProgressiveWidening::Alpha": 0.5,
        P0 + "Mcts::UctStatistic::ProgressiveWidening::K": 0.4,
    }


def hypotheses(num_samples=2000):
    """two IDM hypotheses: headway U[1, 1.5] and U[1.5, 3] (behavior_uct_hypothesis_test.cc:246-252)"""
    return [P.HypothesisIDM(kat_worlds.make_params_hypothesis(1.0, 1.5), num_samples=num_samples),
            P.HypothesisIDM(kat_worlds.make_params_hypothesis(1.5, 3.0), num_samples=num_samples)]


def observed_world(n_others, rel_distance, ego_velocity, velocity_difference,
                   goal_center=(40.0, -1.75), goal_half=4.0):
    """make_test_observed_world(n_others, rel_distance, ego_velocity, velocity_difference, goal)"""
    cfg, scn, pose, alive = kat_worlds.test_world(n_others, rel_distance, ego_velocity,
                                                  velocity_difference, [kat_worlds.make_params_hypothesis(1.0, 3.0)],
                                                  goal_center=goal_center, goal_half=goal_half)
    agents = [P.Agent(id=1 + i, x=float(pose[i, 0, 0]), y=float(pose[i, 0, 1]),
                      theta=float(pose[i, 0, 2]), v=float(pose[i, 0, 3]))
              for i in range(1 + n_others)]
    return P.ObservedWorld(ego_id=1, agents=agents, map=scn, world_time=0.0)
