"""Randomised scenarios: geometry family, lane / agent / hypothesis counts, vehicle shapes,
evaluation and reward parameters, state kind, horizon and sub-step count are drawn per seed; the
CUDA path must reproduce the CPU oracle bit for bit on every one of them.  Complements the five
fixed workloads of test_parity_gpu.py with corners they do not reach (a lone ego, 32 agents with
64 hypotheses on a curved road, chance costs, terminal safe-distance violations, rear safe
distance, alpha != 0, ...)."""
import os

import numpy as np
import pytest

import oracle
import parity
from planner_mcts_b200 import RolloutEngine, _abi, scenarios

pytestmark = pytest.mark.gpu


def random_case(seed):
    rng = np.random.default_rng(1000 + seed)
    family = ["highway", "merge1", "merge2", "curved"][seed % 4]
    if family == "highway":
        length = float(rng.uniform(150.0, 500.0))
        cors, road = scenarios.highway_geometry(length)
        lanes, ego_lane = [0, 1], int(rng.integers(0, 2))
    elif family in ("merge1", "merge2"):
        length = float(rng.uniform(300.0, 600.0))
        n_main = 1 if family == "merge1" else 2
        cors, road, ramp = scenarios.merge_geometry(
            length, n_main_lanes=n_main, ramp_start=float(rng.uniform(60.0, 140.0)),
            taper=float(rng.uniform(60.0, 200.0)), lane_w=float(rng.uniform(3.0, 3.75)))
        lanes, ego_lane = list(range(n_main + 1)), int(rng.integers(0, n_main + 1))
    else:
        n_lanes = int(rng.integers(1, 5))
        cors, road = scenarios.curved_geometry(radius=float(rng.uniform(80.0, 400.0)),
                                               sweep_deg=float(rng.uniform(30.0, 120.0)),
                                               n_lanes=n_lanes, lane_w=float(rng.uniform(3.0, 3.8)))
        lanes, ego_lane, length = list(range(n_lanes)), int(rng.integers(0, n_lanes)), 200.0
    coop = seed % 5 == 3
    A = int(rng.choice([1, 2, 3, 5, 9, 17, 32]))
    n_hw, n_dv = int(rng.choice([1, 2, 8])), int(rng.choice([1, 1, 2, 8]))
    hyps = [] if coop else scenarios.hypothesis_set(
        n_hw, n_dv, headway=(float(rng.uniform(0.3, 1.0)), float(rng.uniform(1.5, 4.0))),
        desired_vel=(float(rng.uniform(5.0, 10.0)), float(rng.uniform(12.0, 20.0))),
        spacing=(1.0, 3.0) if seed % 3 == 0 else (2.0, 2.0),
        max_acc=(1.0, 2.5) if seed % 7 == 0 else (1.7, 1.7))
    kind = _abi.STATE_COOPERATIVE if coop else \
        (_abi.STATE_RISK_CONSTRAINT if seed % 2 else _abi.STATE_HYPOTHESIS)
    cfg = _abi.default_config()
    cfg.state_kind = kind
    cfg.max_rollout_steps = int(rng.integers(1, 40))
    cfg.num_substeps = int(rng.choice([1, 4, 10, 13]))
    cfg.prediction_k = float(rng.uniform(0.2, 0.8))
    cfg.prediction_alpha = float(rng.choice([0.0, 0.0, 0.3]))
    cfg.discount_factor = float(rng.uniform(0.7, 1.0))
    cfg.cooperation_factor = float(rng.uniform(0.0, 1.0))
    cfg.add_safe_dist = int(rng.integers(0, 2))
    cfg.split_safe_dist_collision = int(rng.integers(0, 2)) if kind == _abi.STATE_RISK_CONSTRAINT else 0
    cfg.chance_costs = int(rng.integers(0, 2))
    cfg.static_safe_dist_is_terminal = int(rng.integers(0, 2))
    cfg.dynamic_safe_dist_is_terminal = int(rng.integers(0, 2))
    cfg.out_of_drivable_is_terminal = int(rng.integers(0, 2))
    cfg.dyn_to_rear = int(rng.integers(0, 2))
    cfg.static_lat_safe_dist = float(rng.uniform(0.2, 2.0))
    cfg.static_lon_safe_dist = float(rng.uniform(0.2, 3.0))
    cfg.drivable_lat_safe_dist = float(rng.uniform(0.0, 0.8))
    cfg.drivable_lon_safe_dist = float(rng.uniform(0.0, 0.8))
    cfg.dyn_reaction_time_ego = float(rng.uniform(0.1, 1.0))
    cfg.dyn_reaction_time_others = float(rng.uniform(0.1, 1.5))
    cfg.goal_reward = float(rng.uniform(10.0, 300.0))
    cfg.collision_reward = float(-rng.uniform(100.0, 3000.0))
    cfg.step_reward = float(rng.choice([0.0, -1.0]))
    cfg.safe_dist_violated_reward = float(-rng.uniform(0.0, 50.0))
    cfg.crosstrack_gain = float(rng.uniform(0.3, 2.0))
    # macro-action integration: steering limit on / off, lateral acceleration limits, Euler step
    # (0 = num_substeps steps; 0.02 = the reference default)
    cfg.limit_steering_rate = int(rng.integers(0, 2))
    cfg.lat_acc_max = float(rng.uniform(0.5, 6.0))
    cfg.lat_acc_min = float(-rng.uniform(0.5, 6.0))
    cfg.ego_integration_dt = float(rng.choice([0.0, 0.02, 0.05, 0.07]))
    goal = scenarios.frenet_goal(int(rng.integers(0, len(cors))), max_lat=(0.5, 0.7),
                                 max_angle=(0.1, 0.2), vel=(2.0, 40.0)) \
        if seed % 3 == 1 else scenarios.polygon_goal(
            [(cors[0].px[cors[0].n_pts - 1] - 60.0, cors[0].py[cors[0].n_pts - 1] + 2.0),
             (cors[0].px[cors[0].n_pts - 1] + 5.0, cors[0].py[cors[0].n_pts - 1] + 2.0),
             (cors[0].px[cors[0].n_pts - 1] + 5.0, cors[0].py[cors[0].n_pts - 1] - 2.0),
             (cors[0].px[cors[0].n_pts - 1] - 60.0, cors[0].py[cors[0].n_pts - 1] - 2.0)])
    shapes = [(float(rng.uniform(3.0, 8.0)), float(rng.uniform(1.6, 2.5)), float(rng.uniform(0.8, 2.0)))
              for _ in range(A)]
    acts = scenarios.ego_actions(acc_inputs=tuple(rng.choice([0.0, 1.0, 4.0, -1.0, -8.0, 2.0, -3.0],
                                                             size=int(rng.integers(1, 6)),
                                                             replace=False).tolist()),
                                 lane_changes=bool(rng.integers(0, 2)))
    # cooperative state: some agents pursue their own goal (mcts_state_cooperative.cpp:76-84)
    agent_goals = {}
    if coop and A > 1 and seed % 2 == 1:
        pool = [scenarios.frenet_goal(int(rng.integers(0, len(cors))), max_lat=(0.4, 0.9),
                                      max_angle=(0.1, 0.3), vel=(0.0, 25.0)),
                scenarios.polygon_goal([(40.0, 10.0), (400.0, 10.0), (400.0, -12.0), (40.0, -12.0)]),
                scenarios.polygon_goal([(0.0, 1.0), (90.0, 1.0), (90.0, -5.0), (0.0, -5.0)])]
        for a in range(1, A):
            k = int(rng.integers(0, 4))
            if k:
                agent_goals[a] = pool[k - 1]
    scn = scenarios.build_scenario(cors, road, A, hyps, acts, goal, shapes=shapes,
                                   agent_goals=agent_goals)
    lane_of, s_of = scenarios.place_agents(A, lanes, rng, ego_lane=ego_lane,
                                           ego_s=float(rng.uniform(20.0, 80.0)))
    return cfg, scn, lane_of, s_of, rng


# BMCTS_FUZZ_SEEDS=N widens the sweep (the round's one-off runs: 1500 seeds + 300 planner worlds, all bit-exact)
@pytest.mark.parametrize("seed", range(int(os.environ.get("BMCTS_FUZZ_SEEDS", "40"))))
def test_random_scenario_parity(seed, monkeypatch):
    # lanes per rollout: the engine's choice, or a forced value on every third seed
    # (then also the kernel variant: plain / ego ahead; the engine's own choice for batches
    # this small is the ego-ahead variant)
    if seed % 3 == 2:
        monkeypatch.setenv("BMCTS_LANE_T", str([1, 2, 4, 8, 16, 32][(seed // 3) % 6]))
        monkeypatch.setenv("BMCTS_LANE_AHEAD", str((seed // 18) % 2))
    else:
        monkeypatch.delenv("BMCTS_LANE_T", raising=False)
        monkeypatch.delenv("BMCTS_LANE_AHEAD", raising=False)
    cfg, scn, lane_of, s_of, rng = random_case(seed)
    n = int(rng.choice([1, 33, 200]))
    pose, alive, hyp = scenarios.sample_leaves(scn, n, lane_of, s_of, rng, s_jitter=6.0,
                                               v_range=(0.0, 20.0), dead_prob=0.1)
    depth = rng.integers(1, 6, size=n).astype(np.uint16)
    eng = RolloutEngine(cfg, scn)
    g = eng.rollout(pose, alive, hyp, depth=depth, seed=77 + seed, stream=seed, first_rollout_id=seed << 20,
                    want_agents=True, want_final=True)
    c = oracle.rollout(cfg, scn, pose, alive, hyp, depth=depth, seed=77 + seed, stream=seed,
                       first_rollout_id=seed << 20, n_threads=8)
    parity.assert_results_match(g, c)
    A = scn.n_agents
    action = rng.integers(0, scn.n_ego_actions, size=(A, n)).astype(np.uint8)
    action[0, rng.random(n) < 0.15] = _abi.ACTION_NONE   # not expanded: rolled out as they are
    draw = rng.integers(0, 2**62, size=(A, n)).astype(np.uint64)
    g = eng.expand_rollout(pose, alive, hyp, action, draw, depth=depth, seed=5 + seed, stream=1,
                           first_rollout_id=9, want_agents=True)
    c = oracle.expand_rollout(cfg, scn, pose, alive, hyp, action, draw, depth=depth, seed=5 + seed,
                              stream=1, first_rollout_id=9)
    parity.assert_steps_match(g, c, expanded=action[0] != _abi.ACTION_NONE)
    parity.assert_results_match(g, c)
    # belief likelihoods (BehaviorHypothesisIDM::GetProbability for every agent x hypothesis):
    # the histograms are counts, so the probabilities are equal exactly
    if scn.n_hypotheses > 0 and A > 1:
        W = min(n, 4)
        obs = rng.uniform(-5.0, 3.0, size=(A, W)).astype(np.float32)
        kw = dict(n_samples=int(rng.choice([100, 1000])), n_buckets=int(rng.choice([10, 100, 1000])),
                  lower=-6.0, upper=float(rng.choice([4.0, 8.0])), seed=3 + seed)
        gl = eng.likelihood(pose[:, :W], alive[:W], obs, **kw)
        cl = oracle.likelihood(cfg, scn, pose[:, :W], alive[:W], obs, **kw)
        assert np.array_equal(gl, cl)


@pytest.mark.parametrize("seed", range(int(os.environ.get("BMCTS_FUZZ_PLANNER_SEEDS", "16"))))
def test_random_world_planner_parity(seed, monkeypatch):
    """the planners on random worlds: the GPU-backed search builds exactly the tree of the
    oracle-backed one (same host code, bit-identical leaf evaluations) whatever the wave size,
    the number of trees and how they are spread over threads"""
    from planner_mcts_b200 import planner as P
    monkeypatch.delenv("BMCTS_LANE_T", raising=False)
    cfg, scn, lane_of, s_of, rng = random_case(seed)
    pose, _, _ = scenarios.sample_leaves(scn, 1, lane_of, s_of, rng, s_jitter=2.0, v_range=(2.0, 15.0))
    A = scn.n_agents
    agents = [P.Agent(id=10 + 3 * a, x=float(pose[a, 0, 0]), y=float(pose[a, 0, 1]),
                      theta=float(pose[a, 0, 2]), v=float(pose[a, 0, 3]),
                      length=float(scn.agent_length[a]), width=float(scn.agent_width[a]),
                      rear=float(scn.agent_rear[a]), last_acc=float(rng.uniform(-1.0, 1.0)))
              for a in range(A)]
    world = P.ObservedWorld(ego_id=10, agents=agents, map=scn, world_time=1.4)
    b = "BehaviorUctBase::"
    trees = int(rng.choice([1, 1, 3, 7]))
    params = {b + "Mcts::MaxNumIterations": int(rng.choice([64, 300])), b + "Mcts::MaxSearchTime": 1e9,
              b + "Mcts::Engine::WaveSize": int(rng.choice([1, 7, 32, 0])),
              b + "Mcts::NumParallelMcts": trees, b + "Mcts::UseMultiThreading": int(rng.integers(0, 2)),
              b + "Mcts::Engine::MaxThreads": int(rng.choice([0, 2])),
              b + "MaxNearestAgents": A - 1, b + "Mcts::RandomHeuristic::MaxNumIterations": 15,
              b + "Mcts::RandomSeed": 1000 + seed,
              b + "Mcts::State::EvaluationParameters::AddSafeDist": int(cfg.add_safe_dist),
              b + "Mcts::State::SplitSafeDistCollision": int(cfg.split_safe_dist_collision),
              b + "Mcts::HypothesisStatistic::CostBasedActionSelection": int(rng.integers(0, 2)),
              b + "Mcts::MaxSearchDepth": int(rng.choice([1000, 1000, 2, 4])),
              b + "Mcts::MaxNumNodes": int(rng.choice([2000, 40])),
              b + "Mcts::UctStatistic::ProgressiveWidening::Alpha": float(rng.choice([0.1, 0.5])),
              b + "EgoBehavior::BehaviorIDMLaneTracking::LimitSteeringRate": int(cfg.limit_steering_rate),
              b + "EgoBehavior::BehaviorMotionPrimitives::IntegrationTimeDelta": float(rng.choice([0.02, 0.05])),
              "BehaviorUctRiskConstraint::DefaultAvailableRisk": [0.1, 0.0]}
    hyps = [P.HypothesisIDM(scn.hypotheses[h], num_samples=300) for h in range(scn.n_hypotheses)]
    if cfg.state_kind == _abi.STATE_COOPERATIVE:
        cls, args = P.BehaviorUCTCooperative, (params,)
    elif cfg.state_kind == _abi.STATE_RISK_CONSTRAINT:
        cls, args = P.BehaviorUCTRiskConstraint, (params, hyps)
    else:
        cls, args = P.BehaviorUCTHypothesis, (params, hyps)
    lib = oracle.planner_lib()
    g = cls(*args)
    c = cls(*args, _lib=lib, _create=lambda kind, out: lib.oracle_planner_create(kind, out))
    for step in range(2):                                   # the second call updates the beliefs
        world.world_time = 1.4 + 0.2 * step
        tg, tc = g.Plan(0.2, world), c.Plan(0.2, world)
        assert np.array_equal(g.root_stats(), c.root_stats()), step
        assert g.last_macro_action == c.last_macro_action and np.array_equal(tg, tc)
        assert g.last_num_nodes == c.last_num_nodes
        if hasattr(g, "current_beliefs"):
            for a in agents[1:]:
                assert np.array_equal(g.current_beliefs(a.id), c.current_beliefs(a.id))
    assert g.engine_launches > 0 and c.engine_launches == 0
