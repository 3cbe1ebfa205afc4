"""Behaviour-space tooling: the reference's own test cases
(/root/reference/bark_mcts/models/behavior/tests/py_behavior_space_test.py:29-250) on
planner_mcts_b200.behavior_space, plus the hand-over of a generated hypothesis set and risk
function to the planners."""
import pickle
from collections import defaultdict

import pytest
import scipy.stats

import oracle
import planner_worlds as W
from planner_mcts_b200 import planner as P
from planner_mcts_b200.behavior_models import BehaviorHypothesisIDM, BehaviorIDMStochastic
from planner_mcts_b200.behavior_space import BehaviorSpace
from planner_mcts_b200.parameters import ParameterServer
from planner_mcts_b200.risk_calculation import HypothesisRegion

IDM = "BehaviorIDMStochastic"


def _boundaries(p):
    return p["BehaviorSpace"]["Definition"]["SpaceBoundaries"][IDM]


def _sampling(p):
    return p["BehaviorSpace"]["Sampling"][IDM]


def test_parameter_server_conventions(tmp_path):
    p = ParameterServer()
    assert p["A"]["B"]["x", "description", 3] == 3          # default stored ...
    assert p["A"]["B"]["x", "description", 7] == 3          # ... and kept
    assert p.getReal("A::B::x", "", 0.0) == 3.0
    assert p.getListFloat("A::C::v", "", [1, 2]) == [1.0, 2.0]
    assert p.AddChild("A::B") is p["A"]["B"]
    assert len(p.AddChild("A::B", delete=True)) == 0
    p["A"]["B"]["y"] = {"nested": 1}
    assert p.flatten() == {"A::B::y::nested": 1, "A::C::v": [1, 2]}
    p.Save(tmp_path / "p.json")
    assert ParameterServer(filename=tmp_path / "p.json").ConvertToDict() == p.ConvertToDict()
    assert pickle.loads(pickle.dumps(p)).ConvertToDict() == p.ConvertToDict()
    q = p.clone()
    q["A"]["B"]["y"]["nested"] = 2
    assert p["A"]["B"]["y"]["nested"] == 1


def test_default_config_sampling(tmp_path):
    """:29-78"""
    param_server = ParameterServer()
    space = BehaviorSpace(param_server)
    sampled, model_type, _, _ = space.sample_behavior_parameters()
    assert model_type == IDM
    BehaviorIDMStochastic(sampled)
    param_server.Save(tmp_path / "defaults.json")

    p = ParameterServer(filename=tmp_path / "defaults.json")
    _boundaries(p)["SpacingDistribution"] = [1.23, 20.2]
    _sampling(p)["SpacingDistribution"]["StdRange"] = [0.2, 0.4]
    _sampling(p)["SpacingDistribution"]["DistributionType"] = "NormalDistribution1D"
    _boundaries(p)["DesiredVelDistribution"] = [4.5, 6.07]
    _sampling(p)["DesiredVelDistribution"]["Width"] = [0.1, 0.3]
    _sampling(p)["DesiredVelDistribution"]["DistributionType"] = "UniformDistribution1D"
    _boundaries(p)["MaxAccDistribution"] = [1.7]
    _sampling(p)["MaxAccDistribution"]["DistributionType"] = "FixedValue"
    _boundaries(p)["ComftBrakingDistribution"] = [20]
    _sampling(p)["ComftBrakingDistribution"]["DistributionType"] = "FixedValue"
    space = BehaviorSpace(p)

    for _ in range(100):
        sampled, model_type, _, _ = space.sample_behavior_parameters()
        model = BehaviorIDMStochastic(sampled)
        s = sampled[IDM]
        assert s["MaxAccDistribution"]["DistributionType", "", ""] == "FixedValue"
        assert s["MaxAccDistribution"]["FixedValue", "", 1.0] == [1.7]
        assert s["ComftBrakingDistribution"]["DistributionType", "", ""] == "FixedValue"
        assert s["ComftBrakingDistribution"]["FixedValue", "", 1.2] == [20]
        assert s["SpacingDistribution"]["DistributionType", "", ""] == "NormalDistribution1D"
        assert 0.2 <= s["SpacingDistribution"]["StdDev", "", 0.1] <= 0.4
        assert s["DesiredVelDistribution"]["DistributionType", "", ""] == "UniformDistribution1D"
        lb = s["DesiredVelDistribution"]["LowerBound", "", 1.0]
        ub = s["DesiredVelDistribution"]["UpperBound", "", 0.2]
        assert lb >= 4.5 and ub <= 6.07 + 1e-12
        assert lb + 0.1 - 1e-12 <= ub <= lb + 0.3 + 1e-12
        assert model.distributions["MaxAccDistribution"] == ("fixed", 1.7)
        assert model.distributions["SpacingDistribution"][0] == "normal"


def test_sampling_draws_differ_with_a_shared_random_state():
    import numpy as np
    space = BehaviorSpace(ParameterServer())
    rs = np.random.RandomState(3)
    a = space.sample_behavior_parameters(rs)[0].flatten()
    b = space.sample_behavior_parameters(rs)[0].flatten()
    assert a != b
    # without one, every call restarts from BehaviorSpace::Sampling::RandomSeed (:25-27)
    assert space.sample_behavior_parameters()[0].flatten() == \
        space.sample_behavior_parameters()[0].flatten()


def test_partition_sampling():
    """:80-135"""
    p = ParameterServer()
    BehaviorSpace(p).sample_behavior_parameters()
    _boundaries(p)["SpacingDistribution"] = [2.5, 3.5]
    _sampling(p)["SpacingDistribution"]["StdRange"] = [0.2, 0.4]
    _sampling(p)["SpacingDistribution"]["DistributionType"] = "NormalDistribution1D"
    _sampling(p)["SpacingDistribution"]["Partitions"] = 4
    _sampling(p)["SpacingDistribution"]["SelectedPartition"] = 2
    _boundaries(p)["DesiredVelDistribution"] = [4.8, 6.0]
    _sampling(p)["DesiredVelDistribution"]["Width"] = [0.05, 0.1]
    _sampling(p)["DesiredVelDistribution"]["DistributionType"] = "UniformDistribution1D"
    _sampling(p)["DesiredVelDistribution"]["Partitions"] = 12
    _sampling(p)["DesiredVelDistribution"]["SelectedPartition"] = 7
    _boundaries(p)["MaxAccDistribution"] = [1.7]
    _sampling(p)["MaxAccDistribution"]["DistributionType"] = "FixedValue"
    _boundaries(p)["ComftBrakingDistribution"] = [20]
    _sampling(p)["ComftBrakingDistribution"]["DistributionType"] = "FixedValue"
    space = BehaviorSpace(p)
    import numpy as np
    rs = np.random.RandomState(11)
    for _ in range(100):
        sampled, _, p_prior, p_sampling = space.sample_behavior_parameters(rs)
        s = sampled[IDM]
        assert s["MaxAccDistribution"]["FixedValue", "", 1.0] == [1.7]
        assert s["ComftBrakingDistribution"]["FixedValue", "", 1.2] == [20]
        assert s["SpacingDistribution"]["DistributionType", "", ""] == "NormalDistribution1D"
        assert 0.2 <= s["SpacingDistribution"]["StdDev", "", 0.5] <= 0.4
        assert 3.0 <= s["SpacingDistribution"]["Mean", "", 0.1] <= 3.25
        lb = s["DesiredVelDistribution"]["LowerBound", "", 1.0]
        ub = s["DesiredVelDistribution"]["UpperBound", "", 1.0]
        assert lb >= 5.5 - 1e-12 and ub <= 5.6 + 1e-12
        assert 0.05 - 1e-12 <= ub - lb <= 0.1 + 1e-12
        assert p_prior >= 0.0 and p_sampling > 0.0


def test_default_config_hypothesis_creation(tmp_path):
    """:137-156"""
    p = ParameterServer()
    space = BehaviorSpace(p)
    hypothesis_set, hypothesis_parameters = space.create_hypothesis_set()
    n = p["BehaviorSpace"]["Hypothesis"]["Partitions"][IDM]["HeadwayDistribution"]
    assert len(hypothesis_set) == n == len(hypothesis_parameters)
    for idx, h in enumerate(hypothesis_set):
        assert h.params.getReal(IDM + "::HeadwayDistribution::LowerBound", "", 0.0) == \
            pytest.approx(3.0 + idx / n, abs=1e-5)
        assert h.params.getReal(IDM + "::HeadwayDistribution::UpperBound", "", 0.0) == \
            pytest.approx(3.0 + (idx + 1) / n, abs=1e-5)
    p.Save(tmp_path / "hyp.json")
    BehaviorSpace(ParameterServer(filename=tmp_path / "hyp.json")).create_hypothesis_set()


def test_cover_hypothesis_creation():
    """:158-180"""
    p = ParameterServer(log_if_default=False)
    r1 = _boundaries(p)["HeadwayDistribution"] = [5.34, 10.0]
    r2 = _boundaries(p)["SpacingDistribution"] = [1.3434, 10.0]
    fixed = _boundaries(p)["DesiredVelDistribution"] = [3.4545]
    hypothesis_set, _ = BehaviorSpace(p).create_cover_hypothesis()
    assert len(hypothesis_set) == 1
    params = hypothesis_set[0].params
    assert params.getReal(IDM + "::HeadwayDistribution::LowerBound", "", 0.0) == pytest.approx(r1[0])
    assert params.getReal(IDM + "::HeadwayDistribution::UpperBound", "", 0.0) == pytest.approx(r1[1])
    assert params.getReal(IDM + "::SpacingDistribution::LowerBound", "", 0.0) == pytest.approx(r2[0])
    assert params.getReal(IDM + "::SpacingDistribution::UpperBound", "", 0.0) == pytest.approx(r2[1])
    assert params.getListFloat(IDM + "::DesiredVelDistribution::FixedValue", "", [0.0])[0] == \
        pytest.approx(fixed[0])
    # the engine-side region of the same hypothesis
    region = hypothesis_set[0].region
    assert (region.lo[0], region.hi[0]) == pytest.approx((5.34, 10.0))
    assert region.lo[3] == region.hi[3] == pytest.approx(3.4545)


def test_multiple_hypothesis_sets_creation():
    """:182-250"""
    p = ParameterServer()
    hyp = p["BehaviorSpace"]["Hypothesis"]["BehaviorHypothesisIDM"]
    hyp["BucketsLowerBound"], hyp["BucketsUpperBound"] = -1.5321, 5.2323
    hyp["NumSamples"], hyp["NumBuckets"] = 100433, 12553
    b = _boundaries(p)
    ranges = {"HeadwayDistribution": [5.3434, 10.14], "SpacingDistribution": [1.0, 2.0],
              "DesiredVelDistribution": [3, 4.5]}
    fixed = {"ComftBrakingDistribution": [1.0], "MaxAccDistribution": [15.3],
             "CoolnessFactorDistribution": [0.99], "YieldingDurationDistribution": [1.0],
             "NoYieldingDurationDistribution": [1.0]}
    for k, v in {**ranges, **fixed}.items():
        b[k] = v
    splits = [2, 8]
    collection = BehaviorSpace(p).create_multiple_hypothesis_sets(splits=splits)
    assert sorted(collection) == splits
    for split, (hypothesis_set, params_sets) in collection.items():
        assert len(hypothesis_set) == split ** 3 == len(params_sets)
        seen = defaultdict(set)
        for h in hypothesis_set:
            params = h.params
            for k, v in fixed.items():
                assert params.getListFloat(f"{IDM}::{k}::FixedValue", "", [2334.0])[0] == \
                    pytest.approx(v[0])
            assert params.getReal("BehaviorHypothesisIDM::BucketsLowerBound", "", 2334.0) == -1.5321
            assert params.getReal("BehaviorHypothesisIDM::BucketsUpperBound", "", 2334.0) == 5.2323
            assert params.getInt("BehaviorHypothesisIDM::NumSamples", "", 2334) == 100433
            assert params.getInt("BehaviorHypothesisIDM::NumBuckets", "", 2334) == 12553
            assert (h.num_samples, h.num_buckets) == (100433, 12553)
            assert (h.buckets_lower_bound, h.buckets_upper_bound) == (-1.5321, 5.2323)
            for k in ranges:
                seen[k].add((params.getReal(f"{IDM}::{k}::LowerBound", "", 0.0),
                             params.getReal(f"{IDM}::{k}::UpperBound", "", 0.0)))
        for k, (lo, hi) in ranges.items():
            parts = sorted(seen[k])
            assert len(parts) == split
            for idx, (plo, phi) in enumerate(parts):
                assert plo == pytest.approx(lo + idx * (hi - lo) / split, abs=1e-9)
                assert phi == pytest.approx(lo + (idx + 1) * (hi - lo) / split, abs=1e-9)
        # all combinations are distinct, and a hypothesis survives pickling with its parameters
        assert len({bytes(h.region) for h in hypothesis_set}) == split ** 3
        twin = pickle.loads(pickle.dumps(hypothesis_set[-1]))
        assert bytes(twin.region) == bytes(hypothesis_set[-1].region)
        assert twin.num_samples == 100433


def test_prior_knowledge_probability_estimation():
    """:252-297 (skipped in the reference): density of the truncated-normal prior at the sampled
    value and of the uniform proposal over the selected partition"""
    p = ParameterServer()
    for name, v in [("SpacingDistribution", 2), ("MaxAccDistribution", 3),
                    ("ComftBrakingDistribution", 4), ("CoolnessFactorDistribution", 2.5),
                    ("HeadwayDistribution", 1.2), ("YieldingDurationDistribution", 1.0),
                    ("NoYieldingDurationDistribution", 1.0)]:
        _sampling(p)[name]["DistributionType"] = "FixedValue"
        _boundaries(p)[name] = [v]
    _boundaries(p)["DesiredVelDistribution"] = [0.8, 10.8]
    _sampling(p)["DesiredVelDistribution"]["DistributionType"] = "FixedValue"
    _sampling(p)["DesiredVelDistribution"]["Partitions"] = 10
    _sampling(p)["DesiredVelDistribution"]["SelectedPartition"] = 1
    mean, std, region = 4.0, 0.5, (1.8, 2.8)
    prior = p["BehaviorSpace"]["Definition"]["PriorKnowledgeFunction"][
        "TruncatedNormalKnowledgeFunctionDefinition"][IDM + "::DesiredVelDistribution"]
    prior["Mean"], prior["Std"] = mean, std
    dist = scipy.stats.truncnorm(a=(region[0] - mean) / std, b=(region[1] - mean) / std, loc=mean,
                                 scale=std)
    space = BehaviorSpace(p)
    import numpy as np
    rs = np.random.RandomState(5)
    for _ in range(50):
        sampled, _, p_prior, p_sampling = space.sample_behavior_parameters(rs)
        value = sampled[IDM]["DesiredVelDistribution"]["FixedValue", "", 1000.0]
        assert region[0] <= value <= region[1]
        assert p_prior == pytest.approx(float(dist.pdf(value)), abs=1e-4)
        assert p_sampling == pytest.approx(1.0, abs=1e-4)


def _one_dimensional_space(n_partitions_integration=200):
    p = ParameterServer()
    b = _boundaries(p)
    b["HeadwayDistribution"] = [0.5, 3.5]
    for name, v in [("SpacingDistribution", 2.0), ("MaxAccDistribution", 1.7),
                    ("DesiredVelDistribution", 15.0), ("ComftBrakingDistribution", 1.67),
                    ("CoolnessFactorDistribution", 0.0), ("YieldingDurationDistribution", 1.0),
                    ("NoYieldingDurationDistribution", 1.0)]:
        b[name] = [v]
    d = p["BehaviorSpace"]["Definition"]
    d["PriorKnowledgeFunction"]["FunctionDefinition"] = "UniformKnowledgeFunctionDefinition"
    d["PriorKnowledgeFunction"]["PriorKnowledgeFunction"]["NumPartitionsIntegration"] = \
        n_partitions_integration
    return p


def test_scenario_risk_function_of_a_space():
    """uniform prior 1/3 on [0.5, 3.5], template x + 1 (defaults a = b = 1):
    normalisation constant = 1 / integral(prior * template) = 1 / ((1/3) * 9) = 1/3 ... the
    reference's constant is 1 / sum_cells(prior_integral * template_integral / cell_area)"""
    space = BehaviorSpace(_one_dimensional_space())
    risk = space.get_scenario_risk_function()
    assert risk is space.get_scenario_risk_function()
    template_integral = 0.5 * (3.5 ** 2 - 0.5 ** 2) + 3.0
    assert risk.normalization_constant == pytest.approx(1.0 / (template_integral / 3.0), rel=1e-6)
    hypotheses, _ = space.create_hypothesis_set_fixed_split(4)
    assert len(hypotheses) == 4


def test_generated_hypothesis_set_drives_the_planner():
    """hypotheses, risk function and a ParameterServer tree generated with the behaviour space go
    straight into BehaviorUCTRiskConstraint (CPU-oracle backed here)"""
    space = BehaviorSpace(_one_dimensional_space())
    hypotheses, _ = space.create_hypothesis_set_fixed_split(4)
    assert all(isinstance(h, BehaviorHypothesisIDM) for h in hypotheses)
    for h in hypotheses:
        h.num_samples = 2000
    risk = space.get_scenario_risk_function()
    params = ParameterServer()
    for key, val in W.base_params(64).items():
        params[key, "", val]                      # "A::B::key" paths build the tree
    constraint = params["BehaviorUctRiskConstraint"]
    constraint["EstimateScenarioRisk"] = True
    constraint["DefaultAvailableRisk"] = [0.0, 0.0]
    constraint["UpdateScenarioRisk"] = False
    params["BehaviorUctBase"]["Mcts"]["State"]["SplitSafeDistCollision"] = True
    lib = oracle.planner_lib()
    planner = P.BehaviorUCTRiskConstraint(
        params, hypotheses, scenario_risk_function=risk, _lib=lib,
        _create=lambda kind, out: lib.oracle_planner_create(kind, out))
    world = W.observed_world(1, 6.0, 5.0, 1.0, goal_center=(150.0, -1.75))
    traj = planner.Plan(0.2, world)
    assert traj.shape[1] == 5 and planner.last_num_iterations == 64
    integrals = [risk.CalculateIntegralValue(
        HypothesisRegion(h.region, ["BehaviorIDMStochastic::HeadwayDistribution"])) for h in hypotheses]
    beliefs = planner.current_beliefs(2)
    want = sum(beliefs[i] * integrals[i] / 0.75 for i in range(4)) / len(hypotheses)
    assert planner.last_scenario_risk[0] == pytest.approx(want, rel=1e-9)
    assert planner.last_scenario_risk[0] > 0.0
