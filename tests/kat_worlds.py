"""POD restatement of the reference's test fixtures
(/root/reference/bark_mcts/models/behavior/tests/): BARK make_test_world on the
one-road-two-lanes map -- ego at x=3 on the left lane (y=-1.75), other agents
rel_distance bumper-to-bumper ahead, 4 m x 2 m cars with the reference point
1.25 m ahead of the rear edge."""
import numpy as np

from planner_mcts_b200 import _abi, scenarios

FRONT, REAR = 2.75, 1.25


def make_params_hypothesis(headway_lower, headway_upper, acc_lower=-5.0, acc_upper=8.0):
    """tests/behavior_uct_hypothesis_test.cc:60-92 (make_params_hypothesis)"""
    return scenarios.make_hypothesis(headway=(headway_lower, headway_upper), spacing=(0.0, 0.0),
                                     max_acc=(1.0, 1.0), desired_vel=(15.0, 15.0),
                                     comft_braking=(1.0, 1.0), coolness=(0.0, 0.0),
                                     acc_lower=acc_lower, acc_upper=acc_upper, exponent=4)


def test_world(n_others, rel_distance, ego_velocity, velocity_difference, hypotheses,
               goal_center=(40.0, -1.75), goal_half=4.0, acc_inputs=(0, 1, 4, -1, -8),
               state_kind=_abi.STATE_HYPOTHESIS, length=300.0):
    """make_test_observed_world(n_others, rel_distance, ego_velocity, velocity_difference, goal)"""
    cors, road = scenarios.highway_geometry(length)
    gx, gy = goal_center
    # Polygon(Pose(1,1,0), {(-4,4),(-4,4),(4,4),(4,-4),(-4,-4)}) translated to (gx, gy)
    goal = scenarios.polygon_goal([(gx - goal_half, gy + goal_half), (gx + goal_half, gy + goal_half),
                                   (gx + goal_half, gy - goal_half), (gx - goal_half, gy - goal_half)])
    A = 1 + n_others
    scn = scenarios.build_scenario(cors, road, A, hypotheses,
                                   scenarios.ego_actions(acc_inputs, lane_changes=False), goal)
    pose = np.zeros((A, 1, 4), np.float32)
    pos_x, pos_y = 3.0, -1.75
    pose[0, 0] = (pos_x, pos_y, 0.0, ego_velocity)
    rel = rel_distance + FRONT + REAR
    for i in range(n_others):
        pose[1 + i, 0] = (pos_x + rel + 10.0 * i, pos_y, 0.0, ego_velocity - velocity_difference)
    alive = np.array([(1 << A) - 1], np.uint32)
    cfg = _abi.default_config()
    cfg.state_kind = state_kind
    return cfg, scn, pose, alive
