"""Size-independent properties of the CUDA path at BASELINE.json's full batch sizes (the CPU
oracle cannot follow there): run-to-run determinism, independence of the launch geometry
(lanes per rollout, grid size, host-batch chunking), consistency of the result records, and a
random sample of the big batch re-checked against the oracle."""
import numpy as np
import pytest

import oracle
import parity
from planner_mcts_b200 import RolloutEngine, _abi, scenarios

pytestmark = pytest.mark.gpu


def _run(eng, pose, alive, hyp, **kw):
    return eng.rollout(pose, alive, hyp, seed=1000, stream=5, first_rollout_id=1 << 40, **kw)


@pytest.mark.parametrize("name,n", [("c2_merge_sbg", 1 << 20), ("c5_rsbg", 1 << 17),
                                    ("c4_coop", 1 << 17), ("c3_merge_risk", 1 << 17),
                                    ("c1_highway_mdp", 1 << 17), ("x_curved_max", 1 << 16)])
def test_full_size_batch_properties(name, n, monkeypatch):
    cfg, scn, meta = scenarios.workload(name, horizon=20)
    pose, alive, hyp = scenarios.leaves_for(scn, meta, n, seed=21)
    eng = RolloutEngine(cfg, scn)
    monkeypatch.delenv("BMCTS_LANE_T", raising=False)
    monkeypatch.setenv("BMCTS_PIPE_CHUNK", "0")
    a = _run(eng, pose, alive, hyp)["result"]
    # 1. determinism: the same call again
    assert np.array_equal(a, _run(eng, pose, alive, hyp)["result"])
    # 2. launch geometry does not matter: other lanes-per-rollout, tiny grid, chunked upload
    for t in ("1", "2", "4", "16", "32"):
        monkeypatch.setenv("BMCTS_LANE_T", t)
        assert np.array_equal(a, _run(eng, pose, alive, hyp)["result"]), f"T={t}"
    # the "ego ahead" variant (engine.cu: chosen for small batches) on a full-size one: every
    # column refills many times
    monkeypatch.setenv("BMCTS_LANE_AHEAD", "1")
    for t in ("1", "16"):
        monkeypatch.setenv("BMCTS_LANE_T", t)
        assert np.array_equal(a, _run(eng, pose, alive, hyp)["result"]), f"ego ahead, T={t}"
    monkeypatch.delenv("BMCTS_LANE_AHEAD", raising=False)
    monkeypatch.delenv("BMCTS_LANE_T", raising=False)
    monkeypatch.setenv("BMCTS_PIPE_CHUNK", "100000")
    assert np.array_equal(a, _run(eng, pose, alive, hyp)["result"]), "chunked host batch"
    monkeypatch.setenv("BMCTS_PIPE_CHUNK", "0")
    # the exact early-outs of the drivable-area scan (inner rectangles of the road polygon,
    # safe arcs of the centre lines; found at bmcts_set_scenario) change nothing, final poses
    # included
    monkeypatch.setenv("BMCTS_NO_RECTS", "1")
    plain = RolloutEngine(cfg, scn)
    monkeypatch.delenv("BMCTS_NO_RECTS")
    m = min(n, 1 << 17)
    full = _run(eng, pose[:, :m], alive[:m], hyp[:, :m], want_final=True, want_agents=True)
    b = _run(plain, pose[:, :m], alive[:m], hyp[:, :m], want_final=True, want_agents=True)
    for key in ("result", "final_pose", "ret_agents"):
        assert np.array_equal(full[key], b[key]), f"scan-only path: {key}"
    assert np.array_equal(a, _run(plain, pose, alive, hyp)["result"]), "scan-only path"
    # 3. record consistency
    A, H = meta["A"], cfg.max_rollout_steps
    term = (a["flags"] & _abi.F_TERMINAL) != 0
    assert ((a["n_steps"] >= 1) & (a["n_steps"] <= H)).all()
    assert (a["n_steps"][~term] == H).all()              # only a terminal state ends early
    assert (a["agent_steps"] <= A * a["n_steps"]).all()
    assert (a["agent_steps"] >= a["n_steps"]).all()      # the ego is alive at every step start
    np.testing.assert_allclose(a["step_len"], a["n_steps"] * cfg.prediction_k, rtol=1e-5)
    last, any_ = a["flags"] & 0xFF, (a["flags"] >> 8) & 0xFF
    assert ((last & ~any_) == 0).all()                   # OR over steps contains the last step
    assert (np.bitwise_count(a["final_alive"]) <= A).all()
    if name in ("c2_merge_sbg", "c5_rsbg", "c4_coop"):
        assert len(np.unique(a["n_steps"])) > 3
    # 4. a random sample of the batch against the oracle, with the ids it had in the batch
    rng = np.random.default_rng(3)
    for k in np.sort(rng.choice(n, 48, replace=False)):
        c = oracle.rollout(cfg, scn, pose[:, k:k + 1], alive[k:k + 1], hyp[:, k:k + 1], seed=1000,
                           stream=5, first_rollout_id=(1 << 40) + int(k))
        assert np.array_equal(c["result"], a[k:k + 1]), f"rollout {k}"


def test_root_statistics_are_a_sum_over_streams():
    """root-parallel merge = sum over independent seed streams: running G streams in one
    process and adding the statistics equals what G ranks all-reduce (counts exact)"""
    cfg, scn, meta = scenarios.workload("c2_merge_sbg", horizon=20)
    pose, alive, hyp = scenarios.leaves_for(scn, meta, 4096, seed=2)
    eng = RolloutEngine(cfg, scn)
    rets = []
    for stream in range(4):
        r = eng.rollout(pose, alive, hyp, seed=1000, stream=stream)["result"]
        rets.append(r)
    # distinct streams give distinct rollouts of the same leaves ...
    assert not np.array_equal(rets[0]["ret_ego"], rets[1]["ret_ego"])
    # ... whose means agree within Monte-Carlo noise (4096 samples per stream)
    means = np.array([r["ret_ego"].astype(np.float64).mean() for r in rets])
    sems = np.array([r["ret_ego"].astype(np.float64).std() / np.sqrt(len(r)) for r in rets])
    assert np.abs(means - means.mean()).max() < 5 * sems.max()
