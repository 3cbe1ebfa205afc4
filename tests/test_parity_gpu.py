"""Parity of the CUDA path (through the C ABI) against the CPU oracle on identical
sampled hypotheses, actions and RNG draws.  Bar (BASELINE.json north_star): collision /
violation flags and terminal indices bit-exact; returns, costs and final poses within
1e-4 relative (they are in fact bit-exact, which is asserted too)."""
import numpy as np
import pytest

import kat_worlds
import oracle
import parity
from planner_mcts_b200 import RolloutEngine, _abi, scenarios

pytestmark = pytest.mark.gpu

WORKLOADS = ["c1_highway_mdp", "c2_merge_sbg", "c3_merge_risk", "c4_coop", "c5_rsbg"]
# + a quarter circle at the structure limits (4 lanes x 16 points, 32-point road polygon, Frenet
# goal, lane changes): projections over 15 chords, vertices, no inner rectangle
GEOMETRIES = WORKLOADS + ["x_curved_max"]


@pytest.fixture(params=["auto", "1", "2", "4", "8", "16", "32", "1+ahead", "4+ahead", "16+ahead", "32+ahead"],
                autouse=True)
def lanes_per_rollout(request, monkeypatch):
    """every test runs with the engine's own choice of kernel variant (for these small batches:
    the "ego ahead" variant, engine.cu launch_lane), with each forced number of lanes per
    rollout on the plain variant, and with the ego-ahead variant at 1 / 4 / 16 / 32 lanes
    (BMCTS_LANE_T / BMCTS_LANE_AHEAD are read at every launch; the cooperative kernel's ahead
    variant needs a lane per agent and is the plain one otherwise)"""
    if request.param == "auto":
        monkeypatch.delenv("BMCTS_LANE_T", raising=False)
        monkeypatch.delenv("BMCTS_LANE_AHEAD", raising=False)
    else:
        t, _, ahead = request.param.partition("+")
        monkeypatch.setenv("BMCTS_LANE_T", t)
        monkeypatch.setenv("BMCTS_LANE_AHEAD", "1" if ahead else "0")
    return request.param


@pytest.mark.parametrize("name", ["c2_merge_sbg", "c3_merge_risk", "x_curved_max"])
def test_rollout_parity_with_refill(name, monkeypatch):
    """a 2-block grid: every lane group pulls several rollouts from the work counter"""
    monkeypatch.setenv("BMCTS_LANE_GRID", "2")
    monkeypatch.setenv("BMCTS_LANE_W", "2")
    cfg, scn, meta = scenarios.workload(name, horizon=20)
    n = 1500
    rng = np.random.default_rng(17)
    pose, alive, hyp = scenarios.sample_leaves(scn, n, meta["lane_of"], meta["s_of"], rng,
                                               dead_prob=0.05)
    eng = RolloutEngine(cfg, scn)
    g = eng.rollout(pose, alive, hyp, seed=5, stream=2, first_rollout_id=1 << 33,
                    want_agents=True, want_final=True)
    c = oracle.rollout(cfg, scn, pose, alive, hyp, seed=5, stream=2, first_rollout_id=1 << 33,
                       n_threads=8)
    parity.assert_results_match(g, c)
    assert len(set(g["result"]["n_steps"].tolist())) > 3   # rollouts end at different steps


@pytest.mark.parametrize("name", ["c2_merge_sbg", "c4_coop"])
def test_host_batch_pipeline_parity(name, monkeypatch):
    """a host-memory batch cut into chunks (copy / compute pipeline): ragged last chunk,
    depth array, per-agent returns and final poses scattered back into agent-major arrays"""
    monkeypatch.setenv("BMCTS_PIPE_CHUNK", "64")
    cfg, scn, meta = scenarios.workload(name, horizon=12)
    n = 333
    rng = np.random.default_rng(23)
    pose, alive, hyp = scenarios.sample_leaves(scn, n, meta["lane_of"], meta["s_of"], rng,
                                               dead_prob=0.05)
    depth = rng.integers(1, 4, size=n).astype(np.uint16)
    eng = RolloutEngine(cfg, scn)
    g = eng.rollout(pose, alive, hyp, depth=depth, seed=8, stream=1, first_rollout_id=5000,
                    want_agents=True, want_final=True)
    c = oracle.rollout(cfg, scn, pose, alive, hyp, depth=depth, seed=8, stream=1,
                       first_rollout_id=5000, n_threads=8)
    parity.assert_results_match(g, c)
    assert eng.launch_count() == 6   # ceil(333 / 64) chunks
    # the streamed variant: ONE kernel that starts while the first chunk is in flight and
    # picks rollouts up as their chunks arrive
    monkeypatch.setenv("BMCTS_STREAM_CHUNK", "32")
    g = eng.rollout(pose, alive, hyp, depth=depth, seed=8, stream=1, first_rollout_id=5000,
                    want_agents=True, want_final=True)
    parity.assert_results_match(g, c)
    assert eng.launch_count() == 7


@pytest.mark.parametrize("name", ["c2_merge_sbg", "c4_coop", "c5_rsbg"])
def test_unique_state_form_parity(name, monkeypatch):
    """bmcts_leaf_batch.state_index: n rollouts over a table of distinct states (the waves of a
    search revisit few states under fresh hypothesis samples).  Every host path (packed,
    chunked, streamed) and the device path give the results of the expanded batch."""
    cfg, scn, meta = scenarios.workload(name, horizon=12)
    ns, n = 37, 400
    rng = np.random.default_rng(5)
    pose, alive, _ = scenarios.sample_leaves(scn, ns, meta["lane_of"], meta["s_of"], rng,
                                             dead_prob=0.05)
    depth = rng.integers(1, 4, size=ns).astype(np.uint16)
    index = rng.integers(0, ns, size=n).astype(np.uint32)
    hyp = rng.integers(0, max(1, scn.n_hypotheses), size=(scn.n_agents, n)).astype(np.uint8)
    kw = dict(seed=8, stream=1, first_rollout_id=77)
    want = oracle.rollout(cfg, scn, pose[:, index], alive[index], hyp, depth=depth[index],
                          n_threads=8, **kw)
    c = oracle.rollout(cfg, scn, pose, alive, hyp, depth=depth, state_index=index, n_threads=8, **kw)
    parity.assert_results_match(c, want)          # the oracle's own two forms agree
    eng = RolloutEngine(cfg, scn)
    for env in ({}, {"BMCTS_PIPE_CHUNK": "64"}, {"BMCTS_STREAM_CHUNK": "32"}):
        for k in ("BMCTS_PIPE_CHUNK", "BMCTS_STREAM_CHUNK"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        g = eng.rollout(pose, alive, hyp, depth=depth, state_index=index, want_agents=True,
                        want_final=True, **kw)
        parity.assert_results_match(g, want)
    # device-resident buffers
    import torch
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_pose, d_alive, d_hyp = t(pose.transpose(1, 0, 2)), t(alive.view(np.int32)), t(hyp)   # state-major table
    d_depth, d_index = t(depth.view(np.int16)), t(index.view(np.int32))
    d_res = torch.zeros(n, 8, dtype=torch.float32, device=dev)
    eng.rollout_raw(n, _abi.MEM_DEVICE, d_pose, d_alive, d_hyp, d_depth, d_res, n_states=ns,
                    state_index=d_index, **kw)
    eng.sync()
    from planner_mcts_b200 import RESULT_DTYPE
    got = d_res.cpu().numpy().view(RESULT_DTYPE).reshape(-1)
    assert got.tobytes() == want["result"].tobytes()


def test_cooperative_state_with_per_agent_goals():
    """agents of the cooperative state pursuing their own goals (bmcts_scenario.agent_goal /
    extra_goals; mcts_state_cooperative.cpp:76-84): rollouts and steps bit-exact against the oracle"""
    cfg, scn0, meta = scenarios.workload("c4_coop", horizon=20)
    cors, road = scenarios.highway_geometry(500.0)
    goals = {a: [scenarios.polygon_goal([(100.0, 0.0), (200.0, 0.0), (200.0, -6.0), (100.0, -6.0)]),
                 scenarios.frenet_goal(1, max_lat=(0.5, 0.5), max_angle=(0.2, 0.2), vel=(0.0, 9.0)),
                 scenarios.polygon_goal([(60.0, -3.0), (500.0, -3.0), (500.0, -6.0), (60.0, -6.0)])][a % 3]
             for a in range(1, 20, 2)}
    scn = scenarios.build_scenario(cors, road, 20, [], scenarios.ego_actions(), scn0.goal,
                                   agent_goals=goals)
    pose, alive, hyp = scenarios.leaves_for(scn, meta, 96, seed=5)
    eng = RolloutEngine(cfg, scn)
    g = eng.rollout(pose, alive, hyp, seed=11, stream=1, first_rollout_id=3, want_agents=True,
                    want_final=True)
    c = oracle.rollout(cfg, scn, pose, alive, hyp, seed=11, stream=1, first_rollout_id=3, n_threads=8)
    parity.assert_results_match(g, c)
    plain = oracle.rollout(cfg, scn0, pose, alive, hyp, seed=11, stream=1, first_rollout_id=3,
                           n_threads=8)
    assert not np.array_equal(c["ret_agents"], plain["ret_agents"])   # the goals matter
    rng = np.random.default_rng(2)
    action = rng.integers(0, scn.n_ego_actions, size=(20, 96)).astype(np.uint8)
    g = eng.step(pose, alive, hyp, action, seed=4)
    c = oracle.step(cfg, scn, pose, alive, hyp, action, seed=4)
    parity.assert_steps_match(g, c)


@pytest.mark.parametrize("name", GEOMETRIES)
def test_rollout_parity(name):
    cfg, scn, meta = scenarios.workload(name, horizon=20)
    n = 96 if name != "c5_rsbg" else 64
    pose, alive, hyp = scenarios.leaves_for(scn, meta, n, seed=7)
    eng = RolloutEngine(cfg, scn)
    g = eng.rollout(pose, alive, hyp, seed=1000, stream=3, first_rollout_id=12345,
                    want_agents=True, want_final=True)
    c = oracle.rollout(cfg, scn, pose, alive, hyp, seed=1000, stream=3, first_rollout_id=12345)
    parity.assert_results_match(g, c)
    # the batch must exercise both outcomes
    assert (g["result"]["n_steps"] > 0).all()
    assert eng.launch_count() == 1


@pytest.mark.parametrize("name", ["c2_merge_sbg", "c3_merge_risk", "c4_coop", "x_curved_max"])
def test_step_and_expand_parity(name):
    cfg, scn, meta = scenarios.workload(name, horizon=12)
    n = 70   # ragged: not a multiple of the 32-rollout block
    rng = np.random.default_rng(5)
    pose, alive, hyp = scenarios.sample_leaves(scn, n, meta["lane_of"], meta["s_of"], rng,
                                               dead_prob=0.1)
    A = scn.n_agents
    action = rng.integers(0, scn.n_ego_actions, size=(A, n)).astype(np.uint8)
    draw = rng.integers(0, 2**63, size=(A, n)).astype(np.uint64)
    depth = rng.integers(1, 6, size=n).astype(np.uint16)
    eng = RolloutEngine(cfg, scn)
    g = eng.step(pose, alive, hyp, action, draw, depth=depth, seed=99, stream=1)
    c = oracle.step(cfg, scn, pose, alive, hyp, action, draw, depth=depth, seed=99, stream=1)
    parity.assert_steps_match(g, c)
    g = eng.expand_rollout(pose, alive, hyp, action, draw, depth=depth, seed=99, stream=1,
                           first_rollout_id=777, want_agents=True, want_final=True)
    c = oracle.expand_rollout(cfg, scn, pose, alive, hyp, action, draw, depth=depth, seed=99,
                              stream=1, first_rollout_id=777)
    parity.assert_steps_match(g, c)
    parity.assert_results_match(g, c)


@pytest.mark.parametrize("n", [40, 3000])
def test_asynchronous_halves_of_the_fused_call(n):
    """bmcts_expand_rollout_batch_begin / _end on three engines at once (what a host thread does
    for the trees it interleaves): same bytes as the synchronous call, small waves (zero copy)
    and larger ones (packed copies); the engine refuses other work while a batch is pending"""
    cfg, scn, meta = scenarios.workload("c3_merge_risk", horizon=15)
    rng = np.random.default_rng(41)
    A = scn.n_agents
    engines = [RolloutEngine(cfg, scn) for _ in range(3)]
    inputs = []
    for k, eng in enumerate(engines):
        pose, alive, hyp = scenarios.sample_leaves(scn, n, meta["lane_of"], meta["s_of"], rng,
                                                   dead_prob=0.05)
        action = rng.integers(0, scn.n_ego_actions, size=(A, n)).astype(np.uint8)
        draw = rng.integers(0, 2**62, size=(A, n)).astype(np.uint64)
        inputs.append((pose, alive, hyp, action, draw))
        eng.expand_rollout_begin(pose, alive, hyp, action, draw, seed=3, stream=k,
                                 first_rollout_id=100 * k, want_agents=True)
    with pytest.raises(_abi.BmctsError):
        engines[0].rollout(*inputs[0][:3])          # a batch is pending on this engine
    for k, eng in enumerate(engines):
        got = eng.expand_rollout_end()
        pose, alive, hyp, action, draw = inputs[k]
        want = eng.expand_rollout(pose, alive, hyp, action, draw, seed=3, stream=k,
                                  first_rollout_id=100 * k, want_agents=True)
        for key in want:
            assert got[key].tobytes() == want[key].tobytes(), (k, key)
        c = oracle.expand_rollout(cfg, scn, pose, alive, hyp, action, draw, seed=3, stream=k,
                                  first_rollout_id=100 * k)
        parity.assert_steps_match(got, c)
        parity.assert_results_match(got, c)
    assert engines[0]._lib.bmcts_expand_rollout_batch_end(engines[0]._h) == _abi.OK   # nothing pending


def test_prediction_alpha_and_long_horizon():
    cfg, scn, meta = scenarios.workload("c1_highway_mdp", horizon=200)
    cfg.prediction_k, cfg.prediction_alpha = 0.2, 0.3
    pose, alive, hyp = scenarios.leaves_for(scn, meta, 40, seed=3)
    depth = np.arange(1, 41).astype(np.uint16)
    eng = RolloutEngine(cfg, scn)
    g = eng.rollout(pose, alive, hyp, depth=depth, want_agents=True, want_final=True)
    c = oracle.rollout(cfg, scn, pose, alive, hyp, depth=depth)
    parity.assert_results_match(g, c)


def test_edge_cases_empty_single_dead_ego():
    cfg, scn, meta = scenarios.workload("c2_merge_sbg", horizon=10)
    eng = RolloutEngine(cfg, scn)
    A = scn.n_agents
    # empty batch
    g = eng.rollout(np.zeros((A, 0, 4), np.float32), np.zeros(0, np.uint32),
                    np.zeros((A, 0), np.uint8))
    assert g["result"].shape == (0,)
    # single leaf, ego missing -> out_of_map terminal at the first step
    pose, alive, hyp = scenarios.leaves_for(scn, meta, 1, seed=2)
    alive[:] &= np.uint32(0xFFFFFFFE)
    g = eng.rollout(pose, alive, hyp, want_agents=True, want_final=True)
    c = oracle.rollout(cfg, scn, pose, alive, hyp)
    parity.assert_results_match(g, c)
    assert g["result"]["n_steps"][0] == 1
    assert g["result"]["flags"][0] & _abi.F_OUT_OF_MAP
    # only the ego alive
    pose, alive, hyp = scenarios.leaves_for(scn, meta, 33, seed=4)
    alive[:] = 1
    g = eng.rollout(pose, alive, hyp, want_agents=True, want_final=True)
    c = oracle.rollout(cfg, scn, pose, alive, hyp)
    parity.assert_results_match(g, c)


def test_reference_kat_worlds_on_gpu():
    """the reference's own execute KAT (behavior_uct_hypothesis_test.cc:174-211) through the GPU"""
    hyps = [kat_worlds.make_params_hypothesis(1.0, 1.5), kat_worlds.make_params_hypothesis(1.5, 3.0)]
    cfg, scn, pose, alive = kat_worlds.test_world(1, 10.0, 5.0, -4.0, hyps)
    cfg.goal_reward, cfg.collision_reward, cfg.out_of_drivable_reward = 200.0, -2000.0, -2000.0
    cfg.goal_cost, cfg.collision_cost, cfg.out_of_drivable_cost = 10.0, 300.0, 300.0
    eng = RolloutEngine(cfg, scn)
    hyp = np.zeros((2, 1), np.uint8)
    for ego_action, want_r, want_c in ((2, -2000.0, 300.0), (0, 200.0, 10.0)):
        p, a = pose, alive
        for i in range(1000):
            act = np.array([[ego_action], [0]], np.uint8)
            o = eng.step(p, a, hyp, act, np.full((2, 1), 1000 + i, np.uint64),
                         depth=np.array([1 + i], np.uint16))
            p, a = o["next_pose"], o["next_alive"]
            if o["flags"][0] & _abi.F_TERMINAL:
                break
        assert o["flags"][0] & _abi.F_TERMINAL
        assert o["reward"][0, 0] == pytest.approx(want_r, abs=1e-5)
        assert o["cost"][0, 0] == pytest.approx(want_c, abs=1e-5)


def test_likelihood_parity():
    cfg, scn, meta = scenarios.workload("c2_merge_sbg")
    W = 5
    pose, alive, hyp = scenarios.leaves_for(scn, meta, W, seed=11)
    rng = np.random.default_rng(0)
    action = rng.uniform(-3, 2, size=(scn.n_agents, W)).astype(np.float32)
    eng = RolloutEngine(cfg, scn)
    g = eng.likelihood(pose, alive, action, n_samples=2000, n_buckets=100)
    c = oracle.likelihood(cfg, scn, pose, alive, action, n_samples=2000, n_buckets=100)
    assert np.array_equal(g, c)   # counts / n_samples: exact
    assert g.max() > 0


def test_rollouts_are_reproducible_and_schedule_independent():
    cfg, scn, meta = scenarios.workload("c2_merge_sbg", horizon=15)
    pose, alive, hyp = scenarios.leaves_for(scn, meta, 64, seed=9)
    eng = RolloutEngine(cfg, scn)
    a = eng.rollout(pose, alive, hyp, first_rollout_id=100)["result"]
    # the same leaves in two half batches, second half first
    b1 = eng.rollout(pose[:, 32:], alive[32:], hyp[:, 32:], first_rollout_id=132)["result"]
    b0 = eng.rollout(pose[:, :32], alive[:32], hyp[:, :32], first_rollout_id=100)["result"]
    assert np.array_equal(a, np.concatenate([b0, b1]))


def test_error_codes_instead_of_aborts():
    """the reference aborts (LOG(FATAL), mcts_state_hypothesis.cpp:65,130); the C ABI returns
    status codes and keeps the engine usable"""
    import ctypes as C
    from planner_mcts_b200 import BmctsError
    cfg, scn, meta = scenarios.workload("c2_merge_sbg", horizon=5)
    lib = _abi.load_library()
    # an engine without a scenario
    h = C.c_void_p()
    assert lib.bmcts_create(0, C.byref(cfg), C.byref(h)) == _abi.OK
    inb = _abi.LeafBatch(1, _abi.MEM_HOST, None, None, None, None, 0, 0, 0, 0)
    out = _abi.RolloutOut(None, None, None)
    assert lib.bmcts_rollout_batch(h, C.byref(inb), C.byref(out)) == _abi.ERR_NO_SCENARIO
    unfinal = _abi.Scenario.from_buffer_copy(scn)
    unfinal.finalized = 0
    assert lib.bmcts_set_scenario(h, C.byref(unfinal)) == _abi.ERR_INVALID_ARG
    assert lib.bmcts_create(99, C.byref(cfg), C.byref(C.c_void_p())) == _abi.ERR_INVALID_ARG
    lib.bmcts_destroy(h)
    # null arguments / missing hypothesis ids on a ready engine, which stays usable afterwards
    eng = RolloutEngine(cfg, scn)
    pose, alive, hyp = scenarios.leaves_for(scn, meta, 4, seed=1)
    res = np.zeros(4, dtype=np.dtype([("b", np.uint8, 32)]))
    with pytest.raises(BmctsError) as ei:
        eng.rollout_raw(4, _abi.MEM_HOST, pose, alive, None, None, res)   # hyp_id required
    assert ei.value.status == _abi.ERR_INVALID_ARG
    with pytest.raises(BmctsError):
        eng.rollout_raw(4, _abi.MEM_HOST, None, alive, hyp, None, res)
    g = eng.rollout(pose, alive, hyp)
    c = oracle.rollout(cfg, scn, pose, alive, hyp)
    parity.assert_results_match(g, c)
    assert lib.bmcts_status_string(_abi.ERR_NO_DEVICE).decode().startswith("no CUDA device")


def test_elementary_math_is_bit_identical_on_device_and_host():
    """include/bmcts_math.h compiled by nvcc (device) and by gcc (oracle): every elementary
    function agrees bit for bit on a million random arguments and on the special values
    (signed zeros, NaN, infinities, denormals, exact range boundaries)."""
    cfg, scn, meta = scenarios.workload("c1_highway_mdp")
    eng = RolloutEngine(cfg, scn)
    rng = np.random.default_rng(12)
    n = 1 << 20
    special = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 1e-38, 3e38,
                        np.pi, -np.pi, np.pi / 2, np.pi / 4, 0.41421357, 2.4142137, 1e-3, 50.0],
                       np.float32)
    sa, sb = np.meshgrid(special, special)
    ranges = {"sin": (-400, 400), "cos": (-400, 400), "atan2": (-30, 30), "tan_small": (-0.78, 0.78),
              "div": (-1e3, 1e3), "sqrt": (0, 1e6), "min": (-5, 5), "max": (-5, 5),
              "norm_angle": (-100, 100), "fma_aab": (-100, 100), "pow4": (0, 3)}
    for op in _abi.MATH_OPS:
        lo, hi = ranges[op]
        a = np.concatenate([rng.uniform(lo, hi, n).astype(np.float32), sa.ravel()])
        b = np.concatenate([rng.uniform(lo, hi, n).astype(np.float32), sb.ravel()])
        # functions with a bounded domain (bmcts_math.h: sin / cos / angle wrapping up to a few
        # hundred radians, tan on |x| <= pi/4) get the special values clipped into it; division,
        # sqrt, min / max, fma and atan2 take NaN, infinities and denormals as they are
        if op in ("sin", "cos", "norm_angle", "tan_small", "pow4"):
            a[n:] = np.clip(np.nan_to_num(a[n:], nan=0.0), lo, hi)
        g, c = eng.math_probe(op, a, b), oracle.math_probe(op, a, b)
        same = (g.view(np.uint32) == c.view(np.uint32)) | (np.isnan(g) & np.isnan(c))
        bad = np.nonzero(~same)[0]
        assert bad.size == 0, f"{op}: {bad.size} mismatches, e.g. a={a[bad[:3]]} b={b[bad[:3]]} " \
                              f"gpu={g[bad[:3]]} cpu={c[bad[:3]]}"


def test_wave_paths_without_driver_calls_change_nothing(monkeypatch):
    """a small wave from host memory (the planner's case) runs on the pinned blocks directly and
    reports completion through a flag in host memory (engine.cu wait_wave); with the flag off the
    host synchronizes the stream instead; with kernel timing off there are no event records.
    Same results every way, many waves in a row (the flag's sequence number advances, the work
    counter resets itself between launches), and bmcts_last_kernel_ms says when it has nothing."""
    cfg, scn, meta = scenarios.workload("c3_merge_risk", horizon=20)
    rng = np.random.default_rng(5)
    eng = RolloutEngine(cfg, scn)
    ref = None
    for it in range(12):
        n = [7, 40, 130][it % 3]
        pose, alive, hyp = scenarios.sample_leaves(scn, n, meta["lane_of"], meta["s_of"], rng)
        c = oracle.rollout(cfg, scn, pose, alive, hyp, seed=3, stream=it)
        monkeypatch.delenv("BMCTS_NO_DONE_FLAG", raising=False)
        eng.set_kernel_timing(it % 2 == 0)
        g = eng.rollout(pose, alive, hyp, seed=3, stream=it)
        parity.assert_results_match(g, c)
        if it % 2 == 0:
            assert eng.last_kernel_ms() > 0.0
        else:
            with pytest.raises(Exception, match="timing is off"):
                eng.last_kernel_ms()
        monkeypatch.setenv("BMCTS_NO_DONE_FLAG", "1")
        g2 = eng.rollout(pose, alive, hyp, seed=3, stream=it)
        assert np.array_equal(g["result"], g2["result"])


def test_engine_picks_the_latency_variant_for_small_batches(lanes_per_rollout, monkeypatch):
    """batches too small to fill the GPU (MCTS waves) get a lane per agent and the "ahead" variant
    of the kernel, full-size batches the throughput geometry (engine.cu launch_lane); forced
    values are honoured"""
    if lanes_per_rollout != "auto":
        t, _, ahead = lanes_per_rollout.partition("+")
        cfg, scn, meta = scenarios.workload("c2_merge_sbg", horizon=5)
        eng = RolloutEngine(cfg, scn)
        pose, alive, hyp = scenarios.leaves_for(scn, meta, 40, seed=2)
        eng.rollout(pose, alive, hyp)
        assert eng.last_launch_geometry()[0] == int(t)
        assert eng.last_launch_geometry()[2] == bool(ahead)
        return
    for name, agents, t_small, t_big in (("c2_merge_sbg", 10, 16, 1), ("c5_rsbg", 32, 32, 4),
                                         ("c4_coop", 20, 32, 4), ("c1_highway_mdp", 4, 4, 1)):
        cfg, scn, meta = scenarios.workload(name, horizon=5)
        eng = RolloutEngine(cfg, scn)
        pose, alive, hyp = scenarios.leaves_for(scn, meta, 15, seed=2)
        eng.rollout(pose, alive, hyp)
        assert eng.last_launch_geometry()[0] == t_small and eng.last_launch_geometry()[2], name
        pose, alive, hyp = scenarios.leaves_for(scn, meta, 1 << 16, seed=2)
        eng.rollout(pose, alive, hyp)
        assert eng.last_launch_geometry()[0] == t_big and not eng.last_launch_geometry()[2], name
