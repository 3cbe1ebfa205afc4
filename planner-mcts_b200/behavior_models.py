"""Parameter-tree backed behaviour models: what the reference's Python tooling builds with
`eval("BehaviorIDMStochastic(params)")` / `eval("BehaviorHypothesisIDM(params)")`
(hypothesis/behavior_space/behavior_space.py:237-239).  Constructing one reads every parameter
with its default, which is how the behaviour-space tooling discovers the distribution keys of a
model (behavior_space.py:171-188).  Defaults: tests/data/nheuristic_test.json (BehaviorIDMClassic,
BehaviorIDMStochastic, DynamicModel, BehaviorHypothesisIDM blocks).

`BehaviorHypothesisIDM(params)` is accepted by the planners of `planner.py` wherever a
`HypothesisIDM` is: `.region` is the IDM parameter box the GPU engine samples from.
"""
from . import _abi
from .parameters import ParameterServer
from .planner import HypothesisIDM

IDM_CLASSIC_DEFAULTS = [
    ("MinimumSpacing", 2.0), ("DesiredTimeHeadway", 1.5), ("MaxAcceleration", 1.7),
    ("AccelerationLowerBound", -5.0), ("AccelerationUpperBound", 8.0), ("DesiredVelocity", 15.0),
    ("ComfortableBrakingAcceleration", 1.67), ("MinVelocity", 0.0), ("MaxVelocity", 50.0),
    ("Exponent", 4), ("BrakeForLaneEnd", False), ("BrakeForLaneEndEnabledDistance", 60.0),
    ("BrakeForLaneEndDistanceOffset", 15.0), ("NumTrajectoryTimePoints", 11),
    ("CoolnessFactor", 0.0), ("MaxLatDifferenceToBeFront", 0.0),
    ("MaxAngleDifferenceToBeFront", 2.356194490192345),
]
DYNAMIC_MODEL_DEFAULTS = [("LatAccMax", 4.0), ("LatAccMin", -4.0), ("LonAccelerationMax", 4.0),
                          ("LonAccelerationMin", -8.0)]
# distribution children of BehaviorIDMStochastic, in the order BARK reads them; the first six
# are the dimensions of bmcts_hypothesis.lo / .hi (include/bmcts_c.h)
IDM_STOCHASTIC_DISTRIBUTIONS = [
    "HeadwayDistribution", "SpacingDistribution", "MaxAccDistribution", "DesiredVelDistribution",
    "ComftBrakingDistribution", "CoolnessFactorDistribution"]
IDM_INTENTION_DISTRIBUTIONS = ["YieldingDurationDistribution", "NoYieldingDurationDistribution"]


def read_distribution(node):
    """One distribution child -> ("uniform", lo, hi) | ("fixed", v) | ("normal", mean, std);
    missing entries are filled with their defaults like the C++ constructors do."""
    kind = node["DistributionType", "Distribution type", "UniformDistribution1D"]
    if "Uniform" in kind:
        node["RandomSeed", "Seed of the distribution", 1000]
        return ("uniform", float(node["LowerBound", "Lower bound", 0.0]),
                float(node["UpperBound", "Upper bound", 1.0]))
    if "Fixed" in kind:
        value = node["FixedValue", "The fixed value", [1.0]]
        return ("fixed", float(value[0] if isinstance(value, (list, tuple)) else value))
    if "Normal" in kind:
        node["RandomSeed", "Seed of the distribution", 1000]
        return ("normal", float(node["Mean", "Mean", 0.0]), float(node["StdDev", "Std", 1.0]))
    raise ValueError(f"unknown DistributionType {kind!r}")


class BehaviorIDMStochastic:
    """IDM whose six parameters are redrawn from distributions at every Plan()"""

    def __init__(self, params=None):
        self.params = params if params is not None else ParameterServer()
        classic = self.params.AddChild("BehaviorIDMClassic")
        for key, default in IDM_CLASSIC_DEFAULTS:
            classic[key, "", default]
        stochastic = self.params.AddChild("BehaviorIDMStochastic")
        self.distributions = {}
        for name in IDM_STOCHASTIC_DISTRIBUTIONS:
            self.distributions[name] = read_distribution(stochastic.AddChild(name))
        stochastic["UseIntentionMechanism", "", False]
        for name in IDM_INTENTION_DISTRIBUTIONS:
            self.distributions[name] = read_distribution(stochastic.AddChild(name))
        dyn = self.params.AddChild("DynamicModel")
        for key, default in DYNAMIC_MODEL_DEFAULTS:
            dyn[key, "", default]

    def GetParams(self):
        return self.params

    def to_region(self):
        """the parameter box as a bmcts_hypothesis (uniform and fixed-value dimensions only)"""
        h = _abi.Hypothesis()
        for i, name in enumerate(IDM_STOCHASTIC_DISTRIBUTIONS):
            d = self.distributions[name]
            if d[0] == "uniform":
                h.lo[i], h.hi[i] = d[1], d[2]
            elif d[0] == "fixed":
                h.lo[i] = h.hi[i] = d[1]
            else:
                raise ValueError(f"{name}: a hypothesis region needs uniform or fixed-value "
                                 f"distributions, got {d[0]}")
        classic = self.params["BehaviorIDMClassic"]
        h.acc_lower_bound = classic["AccelerationLowerBound", "", -5.0]
        h.acc_upper_bound = classic["AccelerationUpperBound", "", 8.0]
        h.min_velocity = classic["MinVelocity", "", 0.0]
        h.max_velocity = classic["MaxVelocity", "", 50.0]
        h.exponent = int(classic["Exponent", "", 4])
        return h


class BehaviorHypothesisIDM(BehaviorIDMStochastic, HypothesisIDM):
    """hypothesis/idm/hypothesis_idm.cpp:27-38: a BehaviorIDMStochastic + the settings of its
    likelihood histogram"""

    def __init__(self, params=None):
        BehaviorIDMStochastic.__init__(self, params)
        own = self.params.AddChild("BehaviorHypothesisIDM")
        HypothesisIDM.__init__(
            self, region=self.to_region(),
            num_samples=int(own["NumSamples", "Samples of the likelihood histogram", 10000]),
            num_buckets=int(own["NumBuckets", "Buckets of the likelihood histogram", 100]),
            buckets_lower_bound=float(own["BucketsLowerBound", "Lowest acceleration", -10.0]),
            buckets_upper_bound=float(own["BucketsUpperBound", "Highest acceleration", 10.0]))

    def __getstate__(self):
        return {"params": self.params}

    def __setstate__(self, st):
        self.__init__(st["params"])
