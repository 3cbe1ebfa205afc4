"""The oracle (and, on the GPU box, the CUDA path) against things that do NOT come from this
repo's C / CUDA sources:

  * tests/independent_spec.py: a float64 numpy restatement of DESIGN.md section 3 with different
    algorithms for every geometric predicate (winding number, polygon clipping, vectorised
    projection) and its own Philox -- single steps and whole rollouts on all five BASELINE.json
    workloads, the curved road at the structure limits and randomised scenarios;
  * published known answers: Treiber's IDM equilibrium gap, Philox4x32-10 (Random123 vectors live
    in test_oracle_kat.py), and the longitudinal safe distance of Rizaldi / Immler / Althoff
    checked by brute-force simulation of the two braking vehicles;
  * tests/tiny_mcts.py: a small sequential MCTS written from SURVEY.md Appendix B, against the
    host search of planner-mcts_b200/host at WaveSize = 1 (visit counts, values).

Tolerances: FP32 path vs float64 restatement -- poses 2e-4 + 3e-6 |x| per step (positions of a
few hundred metres in FP32), rewards / costs 1e-4 relative, flags EXACT whenever no predicate is
within MARGIN of its threshold (both sides evaluate the same inequality; closer than that FP32 and
float64 may land on different sides, which the test counts and bounds instead)."""
import math

import numpy as np
import pytest

import independent_spec as S
import oracle
from planner_mcts_b200 import _abi, scenarios

WORKLOADS = ["c1_highway_mdp", "c2_merge_sbg", "c3_merge_risk", "c4_coop", "c5_rsbg", "x_curved_max"]
MARGIN = 2e-3   # metres (m/s, rad for the Frenet goal limits)


def _one_step_case(cfg, scn, pose, alive, hyp, action, draw_id, depth, seed, stream, backend):
    """returns (backend output, independent output) of one execute() on one world"""
    sc = S.Scenario(cfg, scn)
    A = scn.n_agents
    out = backend.step(cfg, scn, pose[:, None], np.array([alive], np.uint32), hyp[:, None],
                       action[:, None], draw_id[:, None], depth=np.array([depth], np.uint16),
                       seed=seed, stream=stream)
    w = S.World(sc, pose.astype(np.float64), alive)
    coop = cfg.state_kind == _abi.STATE_COOPERATIVE
    draws = {i: (S.draw(seed, stream, int(draw_id[i]), i, S.RNG_TREE_PARAMS, 0),
                 S.draw(seed, stream, int(draw_id[i]), i, S.RNG_TREE_PARAMS, 1)) for i in range(1, A)}
    nw, f, rw, costs, last, Dt, margins = S.step(w, hyp, action, draws, depth, coop)
    return out, (nw, f, rw, costs, last, min(margins) if margins else math.inf)


class OracleBackend:
    step = staticmethod(oracle.step)
    rollout = staticmethod(oracle.rollout)


def _compare_step(out, ind, A, alive):
    nw, f, rw, costs, last, margin = ind
    safe = margin > MARGIN
    if safe:
        assert int(out["flags"][0]) == f, (hex(int(out["flags"][0])), hex(f), margin)
        assert int(out["next_alive"][0]) == nw.mask()
    if int(out["next_alive"][0]) == nw.mask():
        for a in range(A):
            if not (alive >> a) & 1:
                continue
            got = out["next_pose"][a, 0]
            want = np.array([nw.x[a], nw.y[a], nw.th[a], nw.v[a]])
            d = got - want
            d[2] = (d[2] + math.pi) % (2 * math.pi) - math.pi
            # FP32 positions of a few hundred metres: an ulp is 3e-5 m, ten sub-steps accumulate
            assert np.abs(d).max() < 2e-4 + 3e-6 * np.abs(want[:2]).max(), (a, got, want)
            assert abs(out["last_action"][a, 0] - last[a]) < 2e-4 * max(1.0, abs(last[a]))
    if safe:
        np.testing.assert_allclose(out["reward"][:, 0], rw, rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(out["cost"][:, 0], costs, rtol=1e-4, atol=1e-4)
    return safe


@pytest.mark.parametrize("name", WORKLOADS)
def test_single_steps_against_the_independent_restatement(name):
    cfg, scn, meta = scenarios.workload(name, horizon=20, ego_dt=0.02 if name in ("c1_highway_mdp", "c4_coop") else 0.05)
    rng = np.random.default_rng(4)
    n = 40 if scn.n_agents <= 16 else 16
    pose, alive, hyp = scenarios.sample_leaves(scn, n, meta["lane_of"], meta["s_of"], rng,
                                               s_jitter=6.0, v_range=(0.0, 20.0), dead_prob=0.08)
    A = scn.n_agents
    n_safe = 0
    for k in range(n):
        action = rng.integers(0, scn.n_ego_actions, size=A).astype(np.uint8)
        draw_id = rng.integers(0, 2**62, size=A).astype(np.uint64)
        depth = int(rng.integers(1, 5))
        out, ind = _one_step_case(cfg, scn, pose[:, k], int(alive[k]), hyp[:, k], action, draw_id,
                                  depth, 1000 + k, 3, OracleBackend)
        n_safe += _compare_step(out, ind, A, int(alive[k]))
    assert n_safe >= 0.5 * n   # the exact flag comparison is not vacuous


def _random_cases():
    import test_fuzz_gpu
    return test_fuzz_gpu.random_case


@pytest.mark.parametrize("seed", range(24))
def test_randomised_scenarios_against_the_independent_restatement(seed):
    """the randomised scenarios of test_fuzz_gpu.py (every evaluation / reward / terminal switch,
    geometry families, shapes, steering limit on / off, Euler step): steps and rollouts"""
    cfg, scn, lane_of, s_of, rng = _random_cases()(seed)
    if cfg.prediction_alpha != 0.0:
        cfg.prediction_alpha = 0.25   # float32(0.3) ** depth differs from float64 0.3 in the 8th digit
    n = 6
    pose, alive, hyp = scenarios.sample_leaves(scn, n, lane_of, s_of, rng, s_jitter=6.0,
                                               v_range=(0.0, 20.0), dead_prob=0.1)
    A = scn.n_agents
    for k in range(n):
        action = rng.integers(0, scn.n_ego_actions, size=A).astype(np.uint8)
        draw_id = rng.integers(0, 2**62, size=A).astype(np.uint64)
        out, ind = _one_step_case(cfg, scn, pose[:, k], int(alive[k]), hyp[:, k], action, draw_id,
                                  int(rng.integers(1, 5)), 7 + seed, seed, OracleBackend)
        _compare_step(out, ind, A, int(alive[k]))
    _compare_rollouts(cfg, scn, pose[:, :3], alive[:3], hyp[:, :3], OracleBackend, seed=5 + seed)


def _compare_rollouts(cfg, scn, pose, alive, hyp, backend, seed=1000, stream=2, first_id=99):
    sc = S.Scenario(cfg, scn)
    n = pose.shape[1]
    got = backend.rollout(cfg, scn, pose, alive, hyp, seed=seed, stream=stream, first_rollout_id=first_id)
    n_safe = 0
    for k in range(n):
        ind = S.rollout(sc, pose[:, k], int(alive[k]), hyp[:, k], 1, seed, stream, first_id + k)
        r = got["result"][k]
        if ind["margin"] <= MARGIN:
            continue    # a predicate came within the margin of its threshold on the way
        n_safe += 1
        assert int(r["n_steps"]) == ind["n_steps"], (k, r, ind)
        assert int(r["flags"]) == ind["flags"] and int(r["final_alive"]) == ind["final_alive"]
        assert int(r["agent_steps"]) == ind["agent_steps"]
        for f in ("ret_ego", "cost0", "cost1", "step_len"):
            assert float(r[f]) == pytest.approx(ind[f], rel=1e-4, abs=1e-3), (k, f)
        if "ret_agents" in got:
            np.testing.assert_allclose(got["ret_agents"][:, k], ind["ret_agents"], rtol=1e-4, atol=1e-3)
        if "final_pose" in got:
            d = got["final_pose"][:, k] - ind["final_pose"]
            d[:, 2] = (d[:, 2] + math.pi) % (2 * math.pi) - math.pi
            live = [(ind["final_alive"] >> a) & 1 == 1 for a in range(scn.n_agents)]
            assert np.abs(d[live]).max(initial=0.0) < 5e-3    # 20 steps of FP32 dynamics
    return n_safe


@pytest.mark.parametrize("name", WORKLOADS)
def test_rollouts_against_the_independent_restatement(name):
    cfg, scn, meta = scenarios.workload(name, horizon=20)
    # lane width 3.5 = 1 + 1.5 + 1: the default envelope of a car on its lane centre TOUCHES the
    # cars of the neighbouring lane exactly (margin 0 at every step); 1.4 keeps the predicate off
    # its threshold so that the flags can be compared
    cfg.static_lat_safe_dist = 1.4
    n = 12 if scn.n_agents <= 16 else 6
    pose, alive, hyp = scenarios.leaves_for(scn, meta, n, seed=21)
    assert _compare_rollouts(cfg, scn, pose, alive, hyp, OracleBackend) >= max(2, n // 4)


# ------------------------------------------------------------------ published known answers
def test_idm_equilibrium_gap_of_treiber_et_al():
    """Treiber, Hennecke, Helbing (2000), eq. (6): following a leader at the same speed, the IDM
    acceleration vanishes at the gap s_e(v) = (s0 + v T) / sqrt(1 - (v / v0)^delta)."""
    T, s0, v0, a, b, delta = 1.4, 2.0, 20.0, 1.2, 1.6, 4
    hyp = scenarios.make_hypothesis(headway=(T, T), spacing=(s0, s0), max_acc=(a, a),
                                    desired_vel=(v0, v0), comft_braking=(b, b), exponent=delta)
    cors, road = scenarios.highway_geometry(400.0)
    scn = scenarios.build_scenario(cors, road, 3, [hyp], scenarios.ego_actions((0.0,), False))
    cfg = _abi.default_config()
    cfg.num_substeps = 1
    for v in (5.0, 12.0, 18.0):
        s_e = (s0 + v * T) / math.sqrt(1.0 - (v / v0) ** delta)
        for gap_factor, sign in ((1.0, 0), (0.8, -1), (1.3, +1)):
            gap = s_e * gap_factor
            pose = np.zeros((3, 1, 4), np.float32)
            pose[0, 0] = (10.0, -4.75, 0.0, 5.0)                       # ego, other lane
            pose[1, 0] = (50.0, -1.75, 0.0, v)                         # follower
            pose[2, 0] = (50.0 + gap + 2.75 + 1.25, -1.75, 0.0, v)     # leader, `gap` bumper to bumper
            out = oracle.step(cfg, scn, pose, np.array([7], np.uint32), np.zeros((3, 1), np.uint8),
                              np.zeros((3, 1), np.uint8))
            acc = float(out["last_action"][1, 0])
            if sign == 0:
                assert abs(acc) < 2e-5 * a * 10
            else:
                assert acc * sign > 1e-3
        # free road: a (1 - (v / v0)^delta)
        pose[2, 0, 0] = 5.0    # the other agent is now behind
        out = oracle.step(cfg, scn, pose, np.array([7], np.uint32), np.zeros((3, 1), np.uint8),
                          np.zeros((3, 1), np.uint8))
        assert float(out["last_action"][1, 0]) == pytest.approx(a * (1 - (v / v0) ** delta), rel=1e-5)


def _brute_force_safe(v_front, v_rear, dist, a_rear, a_front, delta, dt=1e-3):
    """simulate: the front vehicle brakes with a_front from t = 0; the rear vehicle keeps its speed
    for `delta` seconds, then brakes with a_rear; safe = they never touch"""
    t, xf, xr, vf, vr = 0.0, dist, 0.0, v_front, v_rear
    m = dist
    while (vf > 0.0 or vr > 0.0) and t < 60.0:
        ar = a_rear if t >= delta else 0.0
        nvf, nvr = max(0.0, vf + a_front * dt), max(0.0, vr + ar * dt)
        xf += 0.5 * (vf + nvf) * dt
        xr += 0.5 * (vr + nvr) * dt
        vf, vr, t = nvf, nvr, t + dt
        m = min(m, xf - xr)
    return m


def test_longitudinal_safe_distance_against_brute_force_simulation():
    """Rizaldi, Immler, Althoff (iFM 2016): every branch of the case split, checked by simulating
    the two vehicles.  Through the oracle: ego behind a slower / faster / braking leader on the
    highway world, dynamic-safe-distance flag of one step with a zero-length prediction... the
    evaluation runs on the POST-step world, so the brute force starts from the oracle's own
    post-step poses."""
    rng = np.random.default_rng(0)
    hyp = scenarios.make_hypothesis(headway=(1.5, 1.5))
    cors, road = scenarios.highway_geometry(600.0)
    scn = scenarios.build_scenario(cors, road, 2, [hyp], scenarios.ego_actions((0.0,), False))
    n_checked = {True: 0, False: 0}
    branches = set()
    for trial in range(400):
        cfg = _abi.default_config()
        cfg.add_safe_dist = 1
        cfg.dyn_to_rear = 0
        cfg.dyn_max_ego_decel = float(rng.uniform(2.0, 8.0))
        cfg.dyn_max_other_decel = float(rng.uniform(2.0, 8.0))
        cfg.dyn_reaction_time_ego = float(rng.uniform(0.1, 1.5))
        cfg.static_lat_safe_dist = cfg.static_lon_safe_dist = 0.0
        v_e, v_o = float(rng.uniform(0.0, 25.0)), float(rng.uniform(0.0, 25.0))
        gap0 = float(rng.uniform(0.5, 80.0))
        pose = np.zeros((2, 1, 4), np.float32)
        pose[0, 0] = (50.0, -1.75, 0.0, v_e)
        pose[1, 0] = (50.0 + gap0 + 4.0, -1.75, 0.0, v_o)
        out = oracle.step(cfg, scn, pose, np.array([3], np.uint32), np.zeros((2, 1), np.uint8),
                          np.zeros((2, 1), np.uint8))
        if int(out["flags"][0]) & (S.F_COLLISION | S.F_OOM):
            continue
        p = out["next_pose"][:, 0].astype(np.float64)
        dist = (p[1, 0] - p[0, 0]) - 4.0
        if dist <= 0.0:
            continue
        m = _brute_force_safe(p[1, 3], p[0, 3], dist, -cfg.dyn_max_ego_decel,
                              -cfg.dyn_max_other_decel, cfg.dyn_reaction_time_ego)
        if abs(m) < 0.05:
            continue      # too close to the boundary for a 1 ms simulation
        flagged = bool(int(out["flags"][0]) & S.F_DYNAMIC)
        assert flagged == (m < 0.0), (trial, m, p, flagged)
        n_checked[flagged] += 1
        t_stop_f = p[1, 3] / cfg.dyn_max_other_decel
        branches.add((cfg.dyn_reaction_time_ego <= t_stop_f, cfg.dyn_max_other_decel < cfg.dyn_max_ego_decel,
                      p[1, 3] < p[0, 3]))
    assert n_checked[True] > 30 and n_checked[False] > 30
    assert len(branches) == 8     # front still moving at reaction time x who brakes harder x who is faster


# ------------------------------------------------------------------ host search vs a tiny sequential MCTS
def _alias_sampler(beliefs, seed):
    """root-belief sampling per agent slot: Vose alias tables over the beliefs, one uniform double
    draw per slot and iteration (HypothesisBeliefTracker::SlotSampler)"""
    import tiny_mcts
    gen = tiny_mcts.Mt19937(seed)
    tables = []
    for w in beliefs:
        n = len(w)
        tot = sum(max(0.0, x) for x in w)
        prob, alias = [1.0] * n, list(range(n))
        scaled = [max(0.0, x) * n / tot if tot > 0 else 1.0 for x in w]
        small = [i for i in range(n) if scaled[i] < 1.0]
        large = [i for i in range(n) if not scaled[i] < 1.0]
        while small and large:
            lo, hi = small.pop(), large[-1]
            prob[lo], alias[lo] = scaled[lo], hi
            scaled[hi] -= 1.0 - scaled[lo]
            if scaled[hi] < 1.0:
                large.pop()
                small.append(hi)
        tables.append((prob, alias))

    def sample():
        hyp = [0]
        for prob, alias in tables:
            m = gen.next() * len(prob)        # one 32-bit draw: column and fraction
            col, frac = m >> 32, (m & 0xFFFFFFFF) / 4294967296.0
            hyp.append(col if frac < prob[col] else alias[col])
        return hyp
    return sample


@pytest.mark.parametrize("cost_based,hyp_based", [(1, 1), (0, 1), (0, 0)])
def test_host_search_equals_a_tiny_sequential_mcts(cost_based, hyp_based):
    """planner-mcts_b200/host (Mcts, UctStatistic, HypothesisStatistic, belief sampling) at
    WaveSize = 1 against tests/tiny_mcts.py: same visit counts and action values at the root"""
    import kat_worlds
    import planner_worlds as W
    import tiny_mcts
    from planner_mcts_b200 import planner as P
    iterations, n_others = 400, 2
    b = W.P0
    params = W.base_params(iterations)
    params.update({b + "Mcts::Engine::WaveSize": 1, b + "Mcts::MaxNumNodes": 100000,
                   b + "Mcts::RandomHeuristic::MaxNumIterations": 12,
                   b + "Mcts::HypothesisStatistic::CostBasedActionSelection": cost_based,
                   b + "Mcts::HypothesisStatistic::ProgressiveWidening::HypothesisBased": hyp_based,
                   b + "Mcts::HypothesisStatistic::ProgressiveWidening::Alpha": 0.4,
                   b + "Mcts::HypothesisStatistic::ProgressiveWidening::K": 0.5})
    world = W.observed_world(n_others, 6.0, 6.0, 2.0, goal_center=(60.0, -1.75))
    pl = oracle.make_planner(P.BehaviorUCTHypothesis, params, W.hypotheses())
    pl.Plan(0.2, world)
    counts = np.asarray(pl.last_visit_counts)
    values = pl.last_return_values

    regions = [kat_worlds.make_params_hypothesis(1.0, 1.5), kat_worlds.make_params_hypothesis(1.5, 3.0)]
    cfg, scn, pose, alive = kat_worlds.test_world(n_others, 6.0, 6.0, 2.0, regions,
                                                  goal_center=(60.0, -1.75))
    cfg.goal_reward, cfg.collision_reward, cfg.goal_cost, cfg.collision_cost = 200.0, -2000.0, 10.0, 300.0
    cfg.max_rollout_steps = 12
    A = scn.n_agents

    def evaluate(pose_, alive_, depth, hyp, ego_action, draws, rollout_id):
        action = np.zeros((A, 1), np.uint8)
        action[0, 0] = ego_action
        draw_id = np.zeros((A, 1), np.uint64)
        for s, d in draws.items():
            draw_id[s, 0] = d
        o = oracle.expand_rollout(cfg, scn, pose_, np.array([alive_], np.uint32),
                                  np.array(hyp, np.uint8)[:, None], action, draw_id,
                                  depth=np.array([depth], np.uint16), seed=1000, stream=0,
                                  first_rollout_id=rollout_id)
        r = o["result"][0]
        return {"next_pose": o["next_pose"].copy(), "next_alive": int(o["next_alive"][0]),
                "reward": float(o["reward"][0, 0]), "cost0": float(o["cost"][0, 0]),
                "terminal": bool(int(o["flags"][0]) & S.F_TERMINAL),
                "last_action": [o["last_action"][s, 0].tobytes() for s in range(A)],
                "ret": float(r["ret_ego"]), "rcost0": float(r["cost0"])}

    beliefs = [pl.current_beliefs(1 + s).tolist() for s in range(1, A)]
    tiny = {"seed": 1000,
            "uct": {"lower": -1000.0, "upper": 100.0, "c": 0.7, "gamma": 0.9, "bound_estimation": True,
                    "pw_k": 0.4, "pw_alpha": 0.5},
            "hyp": {"k": 0.5, "alpha": 0.4, "hyp_based": bool(hyp_based), "cost_based": bool(cost_based),
                    "lo": 0.0, "hi": 1.0, "c": 0.7, "gamma": 0.9}}
    root = tiny_mcts.search(pose, int(alive[0]), A, scn.n_ego_actions, 2, tiny, iterations,
                            _alias_sampler(beliefs, 2000), evaluate)
    assert root.ego.count == counts.tolist()
    for a, n in enumerate(root.ego.count):
        if n:
            assert values[a] == pytest.approx(root.ego.value[a], rel=1e-12, abs=1e-9)
    assert sum(counts) == iterations and (counts > 0).sum() >= 4


# ------------------------------------------------------------------ GPU: the same, through the C ABI
class GpuBackend:
    engines = {}

    @classmethod
    def _eng(cls, cfg, scn):
        from planner_mcts_b200 import RolloutEngine
        return RolloutEngine(cfg, scn)

    @classmethod
    def step(cls, cfg, scn, pose, alive, hyp, action, draw_id, depth=None, seed=1000, stream=0):
        return cls._eng(cfg, scn).step(pose, alive, hyp, action, draw_id, depth=depth, seed=seed,
                                       stream=stream)

    @classmethod
    def rollout(cls, cfg, scn, pose, alive, hyp, seed=1000, stream=0, first_rollout_id=0):
        return cls._eng(cfg, scn).rollout(pose, alive, hyp, seed=seed, stream=stream,
                                          first_rollout_id=first_rollout_id, want_agents=True,
                                          want_final=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", WORKLOADS)
def test_cuda_path_against_the_independent_restatement(name):
    cfg, scn, meta = scenarios.workload(name, horizon=20)
    cfg.static_lat_safe_dist = 1.4   # see test_rollouts_against_the_independent_restatement
    n = 12 if scn.n_agents <= 16 else 6
    pose, alive, hyp = scenarios.leaves_for(scn, meta, n, seed=21)
    assert _compare_rollouts(cfg, scn, pose, alive, hyp, GpuBackend) >= max(2, n // 4)
    rng = np.random.default_rng(8)
    A = scn.n_agents
    for k in range(min(n, 6)):
        action = rng.integers(0, scn.n_ego_actions, size=A).astype(np.uint8)
        draw_id = rng.integers(0, 2**62, size=A).astype(np.uint64)
        out, ind = _one_step_case(cfg, scn, pose[:, k], int(alive[k]), hyp[:, k], action, draw_id, 2,
                                  1000 + k, 3, GpuBackend)
        _compare_step(out, ind, A, int(alive[k]))
