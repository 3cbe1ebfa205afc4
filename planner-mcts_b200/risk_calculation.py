"""Scenario-risk tooling: Python classes with the names and methods of
the reference's pybind module `bark.core.models.behavior.risk_calculation`
(python_wrapper/python_risk_calculation.cpp:19-123) over the C ABI of include/bmcts_risk_c.h.

    region = PriorKnowledgeRegion({"HeadwayDistribution": (0.5, 3.5)})
    prior = PriorKnowledgeFunction(region, LinearKnowledgeFunctionDefinition(region, params), params)
    risk = prior.CalculateScenarioRiskFunction(template_function)   # -> ScenarioRiskFunction
    planner = BehaviorUCTRiskConstraint(params, hypotheses, scenario_risk_function=risk)

A KnowledgeFunctionDefinition subclass implements CalculateIntegral(region_boundaries) with
region_boundaries a dict name -> (lo, hi), exactly like the reference's Python trampolines.
"""
import ctypes as C

from . import _abi

RISK_SYMBOLS = ["bmcts_risk_region_area", "bmcts_risk_partition", "bmcts_risk_linear_integral",
                "bmcts_risk_normalization_constant"]
_INTEGRAL_FN = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int,
                           C.c_void_p)


def _lib():
    lib = _abi.load_library()
    P = C.POINTER(C.c_double)
    lib.bmcts_risk_region_area.argtypes = [P, P, C.c_int]
    lib.bmcts_risk_region_area.restype = C.c_double
    lib.bmcts_risk_partition.argtypes = [P, P, C.c_int, C.c_int, P, P, C.c_long]
    lib.bmcts_risk_partition.restype = C.c_long
    lib.bmcts_risk_linear_integral.argtypes = [P, P, P, P, C.c_int]
    lib.bmcts_risk_linear_integral.restype = C.c_double
    lib.bmcts_risk_normalization_constant.argtypes = [P, P, C.c_int, C.c_int, _INTEGRAL_FN,
                                                      C.c_void_p, _INTEGRAL_FN, C.c_void_p, P]
    return lib


def _arr(values):
    return (C.c_double * len(values))(*values)


def _param(params, path, default):
    """a parameter from a flat {"A::B": v} dict or from a ParameterServer tree (where reading
    registers the default, like the C++ Params getters)"""
    if hasattr(params, "AddChild"):
        return params[path, "", default]
    return params.get(path, default)


def CalculateRegionBoundariesArea(region_boundaries):
    """knowledge_region.hpp: product of the widths of a {name: (lo, hi)} box"""
    lo = [float(b[0]) for b in region_boundaries.values()]
    hi = [float(b[1]) for b in region_boundaries.values()]
    return _lib().bmcts_risk_region_area(_arr(lo), _arr(hi), len(lo))


class KnowledgeRegion:
    @property
    def definition(self):
        return self.GetDefinition()


class PriorKnowledgeRegion(KnowledgeRegion):
    """prior_knowledge_region.hpp:33-47; dimensions keep their insertion order"""

    def __init__(self, region_definition):
        self._def = {k: (float(v[0]), float(v[1])) for k, v in region_definition.items()}

    def GetDefinition(self):
        return dict(self._def)

    def _bounds(self):
        names = list(self._def)
        return names, [self._def[n][0] for n in names], [self._def[n][1] for n in names]

    def CalculateArea(self):
        names, lo, hi = self._bounds()
        return _lib().bmcts_risk_region_area(_arr(lo), _arr(hi), len(names))

    def Partition(self, num_partitions_per_dimension):
        names, lo, hi = self._bounds()
        d, lib = len(names), _lib()
        cells = lib.bmcts_risk_partition(_arr(lo), _arr(hi), d, int(num_partitions_per_dimension),
                                         None, None, 1 << 26)
        if cells < 0:
            raise ValueError("partition too fine or invalid region")
        out_lo, out_hi = (C.c_double * (cells * d))(), (C.c_double * (cells * d))()
        lib.bmcts_risk_partition(_arr(lo), _arr(hi), d, int(num_partitions_per_dimension), out_lo,
                                 out_hi, cells)
        return [PriorKnowledgeRegion({n: (out_lo[c * d + i], out_hi[c * d + i])
                                      for i, n in enumerate(names)}) for c in range(cells)]


class KnowledgeFunctionDefinition:
    """knowledge_function_definition/knowledge_function_definition.hpp:21-38"""

    def __init__(self, supporting_region, params=None):
        self.supporting_region = supporting_region
        self.params = params if params is not None else {}

    def GetSupportingRegion(self):
        return self.supporting_region

    def CalculateIntegral(self, region_boundaries):
        raise NotImplementedError

    def _callback(self, names):
        def fn(lo, hi, n, _user):
            return float(self.CalculateIntegral({names[i]: (lo[i], hi[i]) for i in range(n)}))
        return _INTEGRAL_FN(fn)


class LinearKnowledgeFunctionDefinition(KnowledgeFunctionDefinition):
    """linear_knowledge_function_definition.hpp:21-58: f = a*x + b per dimension; parameters
    "LinearKnowledgeFunction::<dimension>::a" / "::b", default 1"""

    def __init__(self, supporting_region, params=None):
        super().__init__(supporting_region, params)
        self.function_params = {}
        for name in supporting_region.GetDefinition():
            self.function_params[name] = (
                float(_param(self.params, f"LinearKnowledgeFunction::{name}::a", 1.0)),
                float(_param(self.params, f"LinearKnowledgeFunction::{name}::b", 1.0)))

    def CalculateIntegral(self, region_boundaries):
        names = list(region_boundaries)
        a = [self.function_params[n][0] for n in names]
        b = [self.function_params[n][1] for n in names]
        lo = [region_boundaries[n][0] for n in names]
        hi = [region_boundaries[n][1] for n in names]
        return _lib().bmcts_risk_linear_integral(_arr(a), _arr(b), _arr(lo), _arr(hi), len(names))


class ScenarioRiskFunction:
    """scenario_risk_function.hpp:20-43"""

    def __init__(self, risk_function_definition, normalization_constant):
        self.risk_function_definition = risk_function_definition
        self.normalization_constant = float(normalization_constant)

    def GetIntegralValueTemplateFunction(self, region):
        return self.risk_function_definition.CalculateIntegral(region.GetDefinition())

    def CalculateIntegralValue(self, region):
        return self.normalization_constant * self.GetIntegralValueTemplateFunction(region)


class PriorKnowledgeFunction:
    """prior_knowledge_function.hpp:21-47; "PriorKnowledgeFunction::NumPartitionsIntegration"
    (default 100) cells per dimension"""

    def __init__(self, prior_knowledge_region, knowledge_function, params=None):
        self.prior_knowledge_region = prior_knowledge_region
        self.knowledge_function = knowledge_function
        self.params = params if params is not None else {}
        self.num_partitions_integration = int(
            _param(self.params, "PriorKnowledgeFunction::NumPartitionsIntegration", 100))

    def GetIntegralKnowledeValue(self, knowledge_region):
        return self.knowledge_function.CalculateIntegral(knowledge_region)

    def CalculateScenarioRiskFunction(self, template_scenario_risk_function):
        names, lo, hi = self.prior_knowledge_region._bounds()
        prior_cb = self.knowledge_function._callback(names)
        templ_cb = template_scenario_risk_function._callback(names)
        out = C.c_double()
        st = _lib().bmcts_risk_normalization_constant(
            _arr(lo), _arr(hi), len(names), self.num_partitions_integration, prior_cb, None, templ_cb,
            None, C.byref(out))
        if st != 0:
            raise ValueError("prior knowledge and template risk must be positive on every cell")
        return ScenarioRiskFunction(template_scenario_risk_function, out.value)


# dimension names of a hypothesis region = the IDM parameter they bound
# (BehaviorIDMStochastic::GetParameterRegions; order of bmcts_hypothesis.lo / .hi)
IDM_DIMENSIONS = ["HeadwayDistribution", "SpacingDistribution", "MaxAccDistribution",
                  "DesiredVelDistribution", "ComftBrakingDistribution", "CoolnessFactorDistribution"]


class HypothesisRegion(KnowledgeRegion):
    """the parameter region of a BehaviorHypothesisIDM as a KnowledgeRegion, restricted to the
    dimensions the risk function is defined on"""

    def __init__(self, hypothesis_region, dimension_names):
        self._def = {}
        for name in dimension_names:
            m = next((i for i, d in enumerate(IDM_DIMENSIONS) if d in name or name in d), None)
            if m is None:
                raise KeyError(f"dimension {name!r} is not an IDM parameter distribution")
            self._def[name] = (float(hypothesis_region.lo[m]), float(hypothesis_region.hi[m]))

    def GetDefinition(self):
        return dict(self._def)
