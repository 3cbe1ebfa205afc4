"""Pins the CPU oracle against the known-answer tests the reference holds for the
path, re-expressed on the POD scenario, plus the published
Philox4x32-10 test vectors.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

import kat_worlds
import oracle
from planner_mcts_b200 import _abi


# ------------------------------------------------------------------ Philox

def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32 10 rounds
    assert oracle.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox([0xffffffff] * 4, [0xffffffff] * 2) == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
                         [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


# ---------------------------------------------- flags -> reward / cost / terminal

def _py_reward(cfg, f, dt):
    """mcts_state_base.hpp:107-118 in float64"""
    coll, stat, dyn = bool(f & 1), bool(f & 2), bool(f & 4)
    driv, goal, oom = bool(f & 8), bool(f & 16), bool(f & 32)
    sd = (dyn if cfg.static_safe_dist_is_terminal else (dyn or stat)) * cfg.safe_dist_violated_reward * dt
    col = (coll or (stat if cfg.static_safe_dist_is_terminal else False)) * cfg.collision_reward + \
        (driv or oom) * cfg.out_of_drivable_reward
    return max(col, cfg.collision_reward) + (sd if cfg.add_safe_dist else 0.0) + \
        goal * cfg.goal_reward + cfg.step_reward


def _py_costs(cfg, f, dt):
    """mcts_state_base.hpp:120-142 in float64"""
    coll, stat, dyn = bool(f & 1), bool(f & 2), bool(f & 4)
    driv, goal, oom = bool(f & 8), bool(f & 16), bool(f & 32)
    sdc = (dyn if cfg.static_safe_dist_is_terminal else (dyn or stat)) * cfg.safe_dist_violated_cost * dt
    tot = (coll or (stat if cfg.static_safe_dist_is_terminal else False) or
           (dyn if cfg.dynamic_safe_dist_is_terminal else False)) * cfg.collision_cost + \
        (driv or oom) * cfg.out_of_drivable_cost + \
        (sdc if (cfg.add_safe_dist and not cfg.split_safe_dist_collision) else 0.0) + \
        goal * cfg.goal_cost
    if cfg.split_safe_dist_collision:
        return [min(sdc, 1.0) if cfg.chance_costs else sdc, min(tot, cfg.collision_cost) * dt]
    return [min(tot, 1.0) if cfg.chance_costs else tot, 0.0]


def _py_terminal(cfg, f):
    """mcts_state_base.hpp:273-277"""
    coll, stat, dyn = bool(f & 1), bool(f & 2), bool(f & 4)
    driv, goal, oom = bool(f & 8), bool(f & 16), bool(f & 32)
    return (coll or oom or goal) or (driv and cfg.out_of_drivable_is_terminal) or \
        (stat and cfg.static_safe_dist_is_terminal) or (dyn and cfg.dynamic_safe_dist_is_terminal)


@pytest.mark.parametrize("variant", range(16))
def test_reward_cost_terminal_all_flag_combinations(variant):
    lib = oracle.load()
    cfg = _abi.default_config()
    cfg.add_safe_dist = variant & 1
    cfg.split_safe_dist_collision = (variant >> 1) & 1
    cfg.static_safe_dist_is_terminal = (variant >> 2) & 1
    cfg.dynamic_safe_dist_is_terminal = (variant >> 3) & 1
    cfg.out_of_drivable_is_terminal = variant & 1
    cfg.chance_costs = (variant >> 2) & 1
    cfg.out_of_drivable_reward, cfg.out_of_drivable_cost = -300.0, 40.0
    cfg.step_reward = 0.25
    for dt in (0.5, 0.2):
        for f in range(64):
            r = lib.oracle_reward_from_flags(C.byref(cfg), f, dt)
            assert r == pytest.approx(_py_reward(cfg, f, dt), rel=1e-6, abs=1e-6)
            c = (C.c_float * 2)()
            n = lib.oracle_costs_from_flags(C.byref(cfg), f, dt, c)
            assert n == (2 if cfg.split_safe_dist_collision else 1)
            assert list(c) == pytest.approx(_py_costs(cfg, f, dt), rel=1e-6, abs=1e-6)
            assert bool(lib.oracle_terminal_from_flags(C.byref(cfg), f)) == _py_terminal(cfg, f)


def test_prediction_time_span():
    lib = oracle.load()
    cfg = _abi.default_config()
    assert lib.oracle_prediction_time_span(C.byref(cfg), 1) == 0.5
    assert lib.oracle_prediction_time_span(C.byref(cfg), 7) == 0.5  # alpha = 0
    cfg.prediction_k, cfg.prediction_alpha = 0.2, 0.5
    assert lib.oracle_prediction_time_span(C.byref(cfg), 4) == pytest.approx(0.4, rel=1e-6)


# ----------------------------------------- hypothesis_mcts_state.execute KAT

def _hyp_state_world(state_kind=_abi.STATE_HYPOTHESIS):
    hyps = [kat_worlds.make_params_hypothesis(1.0, 1.5), kat_worlds.make_params_hypothesis(1.5, 3.0)]
    cfg, scn, pose, alive = kat_worlds.test_world(1, 10.0, 5.0, -4.0, hyps, state_kind=state_kind)
    # tests/behavior_uct_hypothesis_test.cc:106-111
    cfg.goal_reward, cfg.collision_reward, cfg.out_of_drivable_reward = 200.0, -2000.0, -2000.0
    cfg.goal_cost, cfg.collision_cost, cfg.out_of_drivable_cost = 10.0, 300.0, 300.0
    return cfg, scn, pose, alive


def _run_until_terminal(cfg, scn, pose, alive, ego_action, other_action=0, max_steps=1000):
    A = scn.n_agents
    hyp = np.zeros((A, 1), np.uint8)
    depth = 1
    o = None
    for i in range(max_steps):
        action = np.zeros((A, 1), np.uint8)
        action[0, 0] = ego_action
        action[1:, 0] = other_action
        draw = np.full((A, 1), 1000 + i, np.uint64)
        o = oracle.step(cfg, scn, pose, alive, hyp, action, draw, depth=np.array([depth], np.uint16))
        pose, alive = o["next_pose"], o["next_alive"]
        depth += 1
        if o["flags"][0] & _abi.F_TERMINAL:
            return o, i + 1
    return o, max_steps


@pytest.mark.parametrize("state_kind", [_abi.STATE_HYPOTHESIS, _abi.STATE_RISK_CONSTRAINT])
def test_hypothesis_state_execute_kat(state_kind):
    """tests/behavior_uct_hypothesis_test.cc:174-211 and
    tests/behavior_uct_risk_constraint_test.cc:174-215"""
    cfg, scn, pose, alive = _hyp_state_world(state_kind)
    hyp = np.zeros((2, 1), np.uint8)
    o = oracle.step(cfg, scn, pose, alive, hyp, np.array([[0], [0]], np.uint8),
                    np.array([[1], [1]], np.uint64))
    assert not (o["flags"][0] & _abi.F_TERMINAL)          # :174
    assert abs(o["reward"][0, 0]) < 1e-5                  # :177
    assert abs(o["cost"][0, 0]) < 1e-3                    # :178
    # the other agent's action is the IDM acceleration it applied (Continuous1DAction)
    assert -5.0 <= o["last_action"][1, 0] <= 8.0
    # repeating ego action 2 (acc +4) leads to a collision with the other agent (:185-198)
    o, steps = _run_until_terminal(cfg, scn, pose, alive, ego_action=2)
    assert o["flags"][0] & _abi.F_TERMINAL and o["flags"][0] & _abi.F_COLLISION_OTHER
    assert o["reward"][0, 0] == pytest.approx(-2000.0, abs=1e-5)
    assert o["cost"][0, 0] == pytest.approx(300.0, abs=1e-5)
    # repeating ego action 0 reaches the goal (:202-211)
    o, steps = _run_until_terminal(cfg, scn, pose, alive, ego_action=0)
    assert o["flags"][0] & _abi.F_GOAL_REACHED
    assert o["reward"][0, 0] == pytest.approx(200.0, abs=1e-5)
    assert o["cost"][0, 0] == pytest.approx(10.0, abs=1e-5)


def test_cooperative_state_execute_kat():
    """tests/behavior_uct_cooperative_test.cc:60-145, cooperation factor 0.75"""
    cfg, scn, pose, alive = kat_worlds.test_world(1, 10.0, 5.0, -4.0, [],
                                                  state_kind=_abi.STATE_COOPERATIVE)
    cfg.goal_reward, cfg.collision_reward = 200.0, -2000.0
    cfg.goal_cost, cfg.collision_cost = 10.0, 300.0
    cfg.cooperation_factor = 0.75
    o = oracle.step(cfg, scn, pose, alive, None, np.array([[0], [2]], np.uint8))
    assert not (o["flags"][0] & _abi.F_TERMINAL)          # :106
    assert abs(o["reward"][0, 0]) < 1e-5 and abs(o["cost"][0, 0]) < 1e-3   # :108-109
    # crash: ego accelerates with +4, the agent in front keeps its speed (:112-130)
    o, _ = _run_until_terminal(cfg, scn, pose, alive, ego_action=2, other_action=0)
    cf, cr = 0.75, -2000.0
    want = ((1 - cf) * cr + cf * cr) * 0.5
    assert o["flags"][0] & _abi.F_TERMINAL
    assert o["reward"][0, 0] == pytest.approx(want, abs=1e-5)      # -1000
    assert o["cost"][0, 0] == pytest.approx(-want, abs=1e-5)       # +1000
    # goal: ego keeps its speed, the other accelerates away (:133-144)
    o, _ = _run_until_terminal(cfg, scn, pose, alive, ego_action=0, other_action=2)
    want = ((1 - cf) * 200.0 + cf * 0.0) / 2.0
    assert o["reward"][0, 0] == pytest.approx(want, abs=1e-5)      # 25
    assert o["cost"][0, 0] == pytest.approx(-want, abs=1e-5)       # -25


# --------------------------------------------------- hypothesis likelihood KAT

def test_likelihood_free_road_is_deterministic():
    """tests/behavior_hypothesis_test.cc:113-158: without a leader the sampled headway does
    not matter, so the planned action has probability ~1 and another action ~0."""
    hyps = [kat_worlds.make_params_hypothesis(1.0, 3.0)]
    for ego_v in (15.0, 25.0):
        cfg, scn, pose, alive = kat_worlds.test_world(0, 20.0, ego_v, 0.0, hyps)
        planned = 1.0 * (1.0 - (ego_v / 15.0) ** 4)
        planned = max(planned, -5.0)
        p = oracle.likelihood(cfg, scn, pose, alive, np.array([[planned]], np.float32),
                              n_samples=10000, n_buckets=100, lower=-8.0, upper=9.0)
        assert p[0, 0, 0] == pytest.approx(1.0, abs=0.01)
        p = oracle.likelihood(cfg, scn, pose, alive, np.array([[planned + 1.5]], np.float32),
                              n_samples=10000, n_buckets=100, lower=-8.0, upper=9.0)
        assert p[0, 0, 0] == pytest.approx(0.0, abs=0.01)


def test_likelihood_density_kat():
    """tests/behavior_hypothesis_test.cc:160-196: gap = v = v0 = 15, s0 = 0, a = b = 1 gives
    a_idm = -T^2 with T ~ U[1,3], so P(a) is proportional to 1/sqrt(|a|)."""
    hyps = [kat_worlds.make_params_hypothesis(1.0, 3.0, acc_lower=-20.0, acc_upper=20.0)]
    cfg, scn, pose, alive = kat_worlds.test_world(1, 15.0, 15.0, 0.0, hyps)
    # query agent = slot 0 (the one with the leader)
    ratios = []
    for a in (-1.5, -2.0, -7.5):
        act = np.array([[a], [0.0]], np.float32)
        p = oracle.likelihood(cfg, scn, pose, alive, act, n_samples=400000, n_buckets=1000,
                              lower=-22.0, upper=20.0)
        ratios.append(p[0, 0, 0] / (1.0 / np.sqrt(4 * abs(a))))
    assert ratios[1] == pytest.approx(ratios[2], abs=0.02 * ratios[1] + 2e-3)
    assert ratios[1] == pytest.approx(ratios[0], abs=0.02 * ratios[1] + 2e-3)
    for a in (-9.1, -0.9):   # outside the support of -T^2 for T in [1, 3]
        act = np.array([[a], [0.0]], np.float32)
        p = oracle.likelihood(cfg, scn, pose, alive, act, n_samples=20000, n_buckets=1000,
                              lower=-22.0, upper=20.0)
        assert p[0, 0, 0] == pytest.approx(0.0, abs=0.01)
    # out-of-bucket action has probability exactly 0 (behavior_uct_hypothesis_test.cc:170-171)
    p = oracle.likelihood(cfg, scn, pose, alive, np.array([[20.5], [0.0]], np.float32),
                          n_samples=200, n_buckets=100, lower=-22.0, upper=20.0)
    assert p[0, 0, 0] == 0.0


def test_idm_acceleration_kat_through_step():
    """a_idm = -T^2 at gap = v = v0 (behavior_hypothesis_test.cc:190-191): with T fixed at
    1.5 the first sub-step acceleration is -2.25; a one-sub-step config exposes it."""
    hyps = [kat_worlds.make_params_hypothesis(1.5, 1.5, acc_lower=-20.0, acc_upper=20.0)]
    cfg, scn, pose, alive = kat_worlds.test_world(1, 15.0, 15.0, 0.0, hyps)
    cfg.num_substeps = 1
    # swap roles: make slot 1 the follower by putting the "ego" in front
    pose = pose[::-1].copy()
    o = oracle.step(cfg, scn, pose, alive, np.zeros((2, 1), np.uint8),
                    np.array([[0], [0]], np.uint8), np.array([[5], [5]], np.uint64))
    assert o["last_action"][1, 0] == pytest.approx(-2.25, rel=1e-5)


def test_unique_state_form_of_a_leaf_batch_equals_the_expanded_batch():
    """bmcts_leaf_batch.state_index: a table of distinct states + one
    index per rollout is the same batch as the states written out per rollout"""
    from planner_mcts_b200 import scenarios
    cfg, scn, meta = scenarios.workload("c3_merge_risk", horizon=8)
    rng = np.random.default_rng(3)
    pose, alive, _ = scenarios.sample_leaves(scn, 9, meta["lane_of"], meta["s_of"], rng, dead_prob=0.1)
    depth = rng.integers(1, 4, size=9).astype(np.uint16)
    index = rng.integers(0, 9, size=150).astype(np.uint32)
    hyp = rng.integers(0, scn.n_hypotheses, size=(scn.n_agents, 150)).astype(np.uint8)
    a = oracle.rollout(cfg, scn, pose, alive, hyp, depth=depth, state_index=index, n_threads=3)
    b = oracle.rollout(cfg, scn, pose[:, index], alive[index], hyp, depth=depth[index])
    for k in a:
        assert a[k].tobytes() == b[k].tobytes(), k


def test_cooperative_agents_are_evaluated_against_their_own_goals():
    """mcts_state_cooperative.cpp:76-84: r_j comes from ObservedWorld(predicted_world, j), i.e. from
    agent j's OWN goal.  Ego goal far away, the other agent's goal around where it will be after the
    step: r_ego = 0, r_1 = GoalReward = 100; with cf = 0.2 and n = 2 the mixed rewards are
    (1/2)((1-cf) 0 + cf 100) = 10 for the ego and (1/2)((1-cf) 100 + cf 0) = 40 for the agent."""
    from planner_mcts_b200 import scenarios
    cors, road = scenarios.highway_geometry(300.0)
    far = scenarios.polygon_goal([(250.0, 0.0), (300.0, 0.0), (300.0, -3.5), (250.0, -3.5)])
    near = scenarios.polygon_goal([(20.0, -2.0), (60.0, -2.0), (60.0, -6.0), (20.0, -6.0)])
    pose = np.zeros((2, 1, 4), np.float32)
    pose[0, 0] = (30.0, -1.75, 0.0, 5.0)
    pose[1, 0] = (40.0, -4.75, 0.0, 5.0)
    cfg = _abi.default_config()
    cfg.state_kind = _abi.STATE_COOPERATIVE
    act = np.zeros((2, 1), np.uint8)
    for goals, want in (({}, (0.0, 0.0)), ({1: near}, (10.0, 40.0))):
        scn = scenarios.build_scenario(cors, road, 2, [], scenarios.ego_actions((0.0,), False), far,
                                       agent_goals=goals)
        out = oracle.step(cfg, scn, pose, np.array([3], np.uint32), np.zeros((2, 1), np.uint8), act)
        assert int(out["flags"][0]) == 0                      # the EGO reached nothing
        assert out["reward"][:, 0].tolist() == pytest.approx(list(want))
        assert out["cost"][0, 0] == pytest.approx(-want[0])   # cost[0] = -reward_ego (:110-111)
