"""Scenario-risk tooling: the reference's own known answers
(/root/reference/bark_mcts/models/behavior/tests/py_scenario_risk_function_test.py:23-239) on
planner_mcts_b200.risk_calculation, and the planner's use of the resulting risk function
(behavior_uct_risk_constraint.cpp:51-72)."""
import pytest

import oracle
import planner_worlds as W
from planner_mcts_b200 import planner as P
from planner_mcts_b200.risk_calculation import (HypothesisRegion, KnowledgeFunctionDefinition,
                                                LinearKnowledgeFunctionDefinition,
                                                PriorKnowledgeFunction, PriorKnowledgeRegion)


def test_prior_knowledge_region_partition_1d():
    """:23-36"""
    partitions = PriorKnowledgeRegion({"1DDimensionName": (-2.0, 5.0)}).Partition(100)
    assert len(partitions) == 100
    total = -2.0
    for part in partitions:
        lo, hi = part.definition["1DDimensionName"]
        assert hi - lo == pytest.approx(7.0 / 100.0, abs=1e-4)
        total += hi - lo
    assert total == pytest.approx(5.0)
    assert partitions[0].definition["1DDimensionName"][0] == pytest.approx(-2.0)
    assert partitions[-1].definition["1DDimensionName"][1] == pytest.approx(5.0)


def test_partition_2d_is_the_cartesian_product():
    region = PriorKnowledgeRegion({"a": (0.0, 1.0), "b": (10.0, 14.0)})
    parts = region.Partition(4)
    assert len(parts) == 16
    assert sum(p.CalculateArea() for p in parts) == pytest.approx(region.CalculateArea())
    assert {p.definition["b"] for p in parts} == {(10.0, 11.0), (11.0, 12.0), (12.0, 13.0), (13.0, 14.0)}


class _Fn(KnowledgeFunctionDefinition):
    def __init__(self, region, integral):
        super().__init__(region, {})
        self._integral = integral

    def CalculateIntegral(self, region_boundaries):
        return self._integral(region_boundaries)


def _abs_integral(indef, name="1DDimensionName", split_at_zero=False):
    def f(rb):
        lo, hi = rb[name]
        if split_at_zero and hi > 0 and lo < 0:
            return abs(indef(0.0) - indef(lo)) + abs(indef(hi) - indef(0.0))
        return abs(indef(hi) - indef(lo))
    return f


def test_1d_functions_uniform():
    """:39-78: uniform prior 1/10, uniform template 1/2 on [-3, 7]"""
    region = PriorKnowledgeRegion({"1DDimensionName": (-3.0, 7.0)})
    prior = _Fn(region, lambda rb: 0.1 * (rb["1DDimensionName"][1] - rb["1DDimensionName"][0]))
    templ = _Fn(region, lambda rb: 0.5 * (rb["1DDimensionName"][1] - rb["1DDimensionName"][0]))
    risk = PriorKnowledgeFunction(region, prior, {}).CalculateScenarioRiskFunction(templ)
    indef = lambda x: 0.1 * 0.5 * x  # noqa: E731
    assert risk.normalization_constant == pytest.approx(1.0 / (indef(7.0) - indef(-3.0)), abs=1e-5)


def test_1d_functions_uniform_and_linear():
    """:80-125: template -x + 10"""
    region = PriorKnowledgeRegion({"1DDimensionName": (-3.0, 7.0)})
    indef = lambda x: -0.5 * x ** 2 + 10 * x  # noqa: E731
    prior = _Fn(region, lambda rb: 0.1 * (rb["1DDimensionName"][1] - rb["1DDimensionName"][0]))
    templ = _Fn(region, _abs_integral(indef, split_at_zero=True))
    risk = PriorKnowledgeFunction(region, prior, {}).CalculateScenarioRiskFunction(templ)
    area = abs(indef(0.0) - indef(-3.0)) + abs(indef(7.0) - indef(0.0))
    assert risk.normalization_constant == pytest.approx(1.0 / (0.1 * area), abs=1e-8)


def test_1d_functions_complex():
    """:127-190: quadratic prior, linear template, 1000 cells; 5 decimals as in the reference"""
    region = PriorKnowledgeRegion({"1DDimensionName": (-3.0, 7.0)})
    a1, b1, c1, a2, b2 = 10.0, 3.0, 2.0, -0.5, 1.0

    def sign_aware(indef):
        def f(rb):
            lo, hi = rb["1DDimensionName"]
            i0, i1 = indef(lo), indef(hi)
            if (i0 < 0 < i1) or (i0 > 0 > i1):
                return abs(i0) + abs(i1)
            return abs(i1 - i0)
        return f

    prior = _Fn(region, sign_aware(lambda x: a1 / 3.0 * x ** 3 + b1 * x ** 2 / 2.0 + c1 * x))
    templ = _Fn(region, sign_aware(lambda x: a2 / 2.0 * x ** 2 + b2 * x))
    risk = PriorKnowledgeFunction(region, prior, {"PriorKnowledgeFunction::NumPartitionsIntegration": 1000}
                                  ).CalculateScenarioRiskFunction(templ)
    indef = lambda x: (a1 * a2 * x ** 4 / 4 + a1 * b2 * x ** 3 / 3 + b1 * a2 * x ** 3 / 3  # noqa: E731
                       + b1 * b2 * x ** 2 / 2 + c1 * a2 * x ** 2 / 2 + c1 * b2 * x)
    area = abs(indef(0.0) - indef(-3.0)) + abs(indef(7.0) - indef(0.0))
    assert risk.normalization_constant == pytest.approx(1.0 / area, abs=1e-5)


def test_2d_functions_complex():
    """:192-239: prior 10*x*y, template -0.5*x on [1,5] x [2,8]"""
    region = PriorKnowledgeRegion({"DimensionName1": (1.0, 5.0), "DimensionName2": (2.0, 8.0)})
    a1, a2 = 10.0, -0.5

    def prior_int(rb):
        (x1, x2), (y1, y2) = rb["DimensionName1"], rb["DimensionName2"]
        return abs(a1 / 4 * (x2 ** 2 * y2 ** 2 - x1 ** 2 * y2 ** 2 - x2 ** 2 * y1 ** 2 + x1 ** 2 * y1 ** 2))

    def templ_int(rb):
        (x1, x2), (y1, y2) = rb["DimensionName1"], rb["DimensionName2"]
        indef = lambda x: a2 / 2.0 * x ** 2  # noqa: E731
        return abs(indef(x2) - indef(x1)) * (y2 - y1)

    risk = PriorKnowledgeFunction(region, _Fn(region, prior_int), {}).CalculateScenarioRiskFunction(
        _Fn(region, templ_int))
    indef = lambda x1, x2, y1, y2: a1 * a2 / 6 * (x2 ** 3 * y2 ** 2 - x1 ** 3 * y2 ** 2  # noqa: E731
                                                   - x2 ** 3 * y1 ** 2 + x1 ** 3 * y1 ** 2)
    assert risk.normalization_constant == pytest.approx(1.0 / abs(indef(1.0, 5.0, 2.0, 8.0)), abs=1e-8)


def test_linear_knowledge_function_and_non_positive_cells():
    region = PriorKnowledgeRegion({"x": (0.0, 2.0)})
    lin = LinearKnowledgeFunctionDefinition(region, {"LinearKnowledgeFunction::x::a": 2.0,
                                                     "LinearKnowledgeFunction::x::b": 1.0})
    # integral of 2x + 1 over [0, 2] by the mid-point mean: ((1) + (5)) / 2 * 2 = 6
    assert lin.CalculateIntegral({"x": (0.0, 2.0)}) == pytest.approx(6.0)
    zero = _Fn(region, lambda rb: 0.0)
    with pytest.raises(ValueError):   # the reference aborts (BARK_EXPECT_TRUE), we raise
        PriorKnowledgeFunction(region, lin, {}).CalculateScenarioRiskFunction(zero)


def test_planner_estimates_the_scenario_risk_from_the_risk_function():
    """EstimateScenarioRisk: the first Plan() sets the available risk to
    sum_h mean-belief(h) * integral(h) / area(h) (behavior_uct_risk_constraint.cpp:51-72)"""
    hyps = W.hypotheses()
    region = PriorKnowledgeRegion({"HeadwayDistribution": (1.0, 3.0)})
    prior = LinearKnowledgeFunctionDefinition(region, {"LinearKnowledgeFunction::HeadwayDistribution::a": 0.0,
                                                       "LinearKnowledgeFunction::HeadwayDistribution::b": 0.5})
    templ = LinearKnowledgeFunctionDefinition(region, {"LinearKnowledgeFunction::HeadwayDistribution::a": -0.1,
                                                       "LinearKnowledgeFunction::HeadwayDistribution::b": 0.4})
    risk = PriorKnowledgeFunction(region, prior, {}).CalculateScenarioRiskFunction(templ)
    integrals = [risk.CalculateIntegralValue(HypothesisRegion(h.region, ["HeadwayDistribution"]))
                 for h in hyps]
    assert all(v > 0 for v in integrals)
    params = W.base_params(64)
    params.update({"BehaviorUctRiskConstraint::EstimateScenarioRisk": 1,
                   "BehaviorUctRiskConstraint::DefaultAvailableRisk": [0.0, 0.0],
                   "BehaviorUctRiskConstraint::UpdateScenarioRisk": 0,
                   W.P0 + "Mcts::State::SplitSafeDistCollision": 1})
    lib = oracle.planner_lib()
    pl = P.BehaviorUCTRiskConstraint(params, hyps, scenario_risk_function=risk, _lib=lib,
                                     _create=lambda kind, out: lib.oracle_planner_create(kind, out))
    world = W.observed_world(1, 6.0, 5.0, 1.0, goal_center=(150.0, -1.75))
    pl.Plan(0.2, world)
    beliefs = pl.current_beliefs(2)
    areas = [0.5, 1.5]   # widths of the two headway ranges [1, 1.5] and [1.5, 3]
    want = sum(beliefs[h] * integrals[h] / areas[h] for h in range(2)) / len(hyps)
    assert pl.last_scenario_risk[0] == pytest.approx(want, rel=1e-9)
    assert pl.last_scenario_risk[0] > 0.0
