"""The C-ABI library loads and exports every symbol include/bmcts_c.h declares; the
ctypes mirror has the header's struct sizes; the host-only entry points behave.  No
compute calls (no GPU needed)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
from conftest import has_gpu
from planner_mcts_b200 import _abi, scenarios

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions(header="bmcts_c.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bmcts_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _abi.load_library()
    declared = _declared_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"libbmcts.so does not export {name}"
    assert sorted(_abi.EXPORTED_SYMBOLS) == declared
    assert lib.bmcts_abi_version() == 2
    # the planner ABI (include/bmcts_planner_c.h)
    from planner_mcts_b200 import planner
    declared = [f for f in _declared_functions("bmcts_planner_c.h") if f.startswith("bmcts_planner_")]
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"libbmcts.so does not export {name}"
    assert sorted(planner.PLANNER_SYMBOLS) == sorted(declared)
    # the scenario-risk tooling ABI (include/bmcts_risk_c.h)
    from planner_mcts_b200 import risk_calculation
    declared = [f for f in _declared_functions("bmcts_risk_c.h") if f.startswith("bmcts_risk_")
                and not f.endswith("_fn")]
    for name in declared:
        assert hasattr(lib, name), f"libbmcts.so does not export {name}"
    assert sorted(risk_calculation.RISK_SYMBOLS) == sorted(declared)
    assert sorted(os.listdir(os.path.join(ROOT, "include"))) == [
        "bmcts_c.h", "bmcts_math.h", "bmcts_planner_c.h", "bmcts_risk_c.h"]


def test_struct_sizes_match_header():
    lib = _abi.load_library()
    for i, s in enumerate(_abi.STRUCTS_IN_HEADER_ORDER):
        assert lib.bmcts_struct_size(i) == C.sizeof(s), s.__name__
    assert lib.bmcts_struct_size(99) == 0
    assert C.sizeof(_abi.RolloutResult) == 32
    assert C.sizeof(_abi.Scenario) % 16 == 0   # staged into shared memory with 16-byte copies


def test_config_defaults_are_the_reference_defaults():
    """param_config/mcts_state_parameters_from_param_server.hpp:19-51,
    param_config/mcts_parameters_from_param_server.hpp:45,55-58"""
    c = _abi.default_config()
    assert (c.goal_reward, c.collision_reward, c.safe_dist_violated_reward,
            c.out_of_drivable_reward) == (100.0, -1000.0, pytest.approx(-0.1), 0.0)
    assert (c.goal_cost, c.collision_cost, c.safe_dist_violated_cost,
            c.out_of_drivable_cost) == (-100.0, 1000.0, pytest.approx(0.1), 0.0)
    assert c.cooperation_factor == pytest.approx(0.2) and c.step_reward == 0.0
    assert c.prediction_k == 0.5 and c.prediction_alpha == 0.0
    assert not (c.split_safe_dist_collision or c.chance_costs or c.add_safe_dist)
    assert c.discount_factor == pytest.approx(0.9) and c.max_rollout_steps == 1000
    assert c.num_substeps == 10
    # ego macro-action integration: tests/data/nheuristic_test.json:153-155 (IntegrationTimeDelta
    # 0.019999999552965164 = float32(0.02)), :183-186 (LimitSteeringRate true), :191-196 (LatAcc)
    assert c.limit_steering_rate == 1
    assert (c.lat_acc_max, c.lat_acc_min) == (4.0, -4.0)
    assert c.ego_integration_dt == np.float32(0.019999999552965164)
    assert _abi.load_library().bmcts_config_validate(C.byref(c)) == _abi.OK
    c.delta_max = 1.0   # > pi/4: outside the validity range of bm_tan_small
    assert _abi.load_library().bmcts_config_validate(C.byref(c)) == _abi.ERR_INVALID_ARG


def test_scenario_finalize_geometry_and_validation():
    cfg, scn, meta = scenarios.workload("c2_merge_sbg")
    ramp = scn.corridors[1]
    px, py = np.array(ramp.px[:4], np.float64), np.array(ramp.py[:4], np.float64)
    seg = np.hypot(np.diff(px), np.diff(py))
    np.testing.assert_allclose(ramp.seg_len[:3], seg, rtol=1e-6)
    np.testing.assert_allclose(ramp.s0[:3], np.concatenate([[0], np.cumsum(seg)[:-1]]), rtol=1e-6)
    np.testing.assert_allclose(ramp.angle[:3], np.arctan2(np.diff(py), np.diff(px)), atol=2e-7)
    np.testing.assert_allclose(np.hypot(ramp.tx[:3], ramp.ty[:3]), 1.0, rtol=1e-6)
    assert ramp.length == pytest.approx(seg.sum(), rel=1e-6)
    lib = _abi.load_library()
    bad = _abi.Scenario.from_buffer_copy(scn)
    bad.n_agents = 33
    assert lib.bmcts_scenario_finalize(C.byref(bad)) == _abi.ERR_INVALID_ARG and not bad.finalized
    bad = _abi.Scenario.from_buffer_copy(scn)
    bad.hypotheses[0].hi[5] = 0.5   # coolness blend is out of scope: rejected, not ignored
    assert lib.bmcts_scenario_finalize(C.byref(bad)) == _abi.ERR_UNSUPPORTED
    bad = _abi.Scenario.from_buffer_copy(scn)
    bad.corridors[0].px[1] = bad.corridors[0].px[0]   # zero-length segment
    assert lib.bmcts_scenario_finalize(C.byref(bad)) == _abi.ERR_INVALID_ARG


def test_projection_matches_numpy():
    cfg, scn, meta = scenarios.workload("c2_merge_sbg")
    lib = oracle.load()
    ramp = scn.corridors[1]
    rng = np.random.default_rng(0)
    for _ in range(200):
        s_true = rng.uniform(1.0, ramp.length - 1.0)
        d_true = rng.uniform(-1.0, 1.0)
        x, y, th = scenarios.pose_on_corridor(ramp, s_true)
        px, py = x - d_true * np.sin(th), y + d_true * np.cos(th)
        s, d, seg, inr = C.c_float(), C.c_float(), C.c_int(), C.c_int()
        lib.oracle_project(C.byref(ramp), float(px), float(py), C.byref(s), C.byref(d),
                           C.byref(seg), C.byref(inr))
        assert inr.value == 1
        # near a polyline vertex the nearest point is not the Frenet foot: loose tolerance there
        assert abs(d.value) <= abs(d_true) + 1e-3
        if abs(s.value - s_true) > 0.05:
            knots = np.array(ramp.s0[:3])
            assert np.min(np.abs(knots - s_true)) < 1.5


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_create_fails_loudly_without_a_gpu():
    """no CPU fallback: the product path refuses to run without a CUDA device"""
    from planner_mcts_b200 import BmctsError, RolloutEngine
    cfg, scn, meta = scenarios.workload("c1_highway_mdp")
    with pytest.raises(BmctsError) as ei:
        RolloutEngine(cfg, scn)
    assert ei.value.status == _abi.ERR_NO_DEVICE


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_planner_fails_loudly_without_a_gpu():
    """the BehaviorUCT* planners refuse to plan without the CUDA engine"""
    import planner_worlds as W
    from planner_mcts_b200 import BmctsError, planner as P
    pl = P.BehaviorUCTHypothesis(W.base_params(10), W.hypotheses())
    with pytest.raises(BmctsError):
        pl.Plan(0.2, W.observed_world(1, 5.0, 5.0, 0.0))


def test_product_does_not_reference_the_oracle():
    """only tests/, bench.py and __graft_entry__.py may touch oracle/"""
    pkg = os.path.join(ROOT, "planner-mcts_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".c")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"#\s*include\s*[\"<][^\">]*oracle", text), f
                assert "liboracle" not in text, f
                assert not re.search(r"\boracle_[a-z_0-9]+\s*\(", text) or f.endswith(".cuh"), f
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
    # and the built library has no undefined oracle symbol
    import subprocess
    syms = subprocess.run(["nm", "-D", os.path.join(pkg, "libbmcts.so")], capture_output=True,
                          text=True).stdout
    assert "oracle_" not in syms
