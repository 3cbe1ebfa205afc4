"""Behaviour-space tooling: hypothesis-set generation and sampling of
behaviour parameters over a box of IDM parameter distributions, with the class and method names of
the reference's Python module (hypothesis/behavior_space/behavior_space.py:18-343 and
knowledge_function_definitions.py:18-105), on this package's `ParameterServer`, behaviour models
and `risk_calculation` classes.

    space = BehaviorSpace(params)                           # params["BehaviorSpace"][...]
    hypotheses, hypothesis_params = space.create_hypothesis_set_fixed_split(8)
    planner = BehaviorUCTRiskConstraint(planner_params, hypotheses, space.get_scenario_risk_function())
    model_params, model_type, p_prior, p_sampling = space.sample_behavior_parameters()

Parameter layout under "BehaviorSpace" (same keys as the reference):
  Definition::ModelType                      model the space is defined over (BehaviorIDMStochastic)
  Definition::SpaceBoundaries::<model tree>  every "...Distribution" key is a range [lo, hi] or a
                                             fixed value [v]; other keys are plain values
  Definition::PriorKnowledgeFunction         FunctionDefinition + its parameters
  Definition::ScenarioRiskFunctionFunction   FunctionName + its parameters
  Sampling::<model tree>::<distribution>     DistributionType, Width / StdRange, Partitions, ...
  Hypothesis::{RandomSeed, HypothesisModel, <model block>, Partitions::<model tree>}

Deviation from the reference, on purpose: the truncated-normal / uniform knowledge functions
integrate their density exactly (product of CDF differences); the reference evaluates the pdf at
`hi - lo` (a width, marked "todo" at knowledge_function_definitions.py:35-41), which is zero for
every narrow cell and makes the normalisation constant of the risk function undefined.
"""
import itertools

import numpy as np

from .behavior_models import BehaviorHypothesisIDM, BehaviorIDMStochastic  # noqa: F401
from .parameters import DELIMITER, ParameterServer
from .risk_calculation import (CalculateRegionBoundariesArea, KnowledgeFunctionDefinition,
                               LinearKnowledgeFunctionDefinition, PriorKnowledgeFunction,
                               PriorKnowledgeRegion)

DEFAULT_RANGE = [3.0, 4.0]
MODEL_TYPES = {"BehaviorIDMStochastic": BehaviorIDMStochastic,
               "BehaviorHypothesisIDM": BehaviorHypothesisIDM}


def _is_distribution(key):
    return "Distribution" in key


def _distribution_entries(tree, prefix=()):
    """(path tuple, value) of every "...Distribution" entry of a parameter tree, depth first in
    insertion order"""
    for key, val in tree.store.items():
        if _is_distribution(key):
            yield prefix + (key,), val
        elif isinstance(val, ParameterServer):
            yield from _distribution_entries(val, prefix + (key,))


def _node(tree, path):
    for part in path:
        tree = tree[part]
    return tree


def _partition_of(param_range, partition_num, which):
    width = param_range[1] - param_range[0]
    return [param_range[0] + float(which) * width / partition_num,
            param_range[0] + float(which + 1) * width / partition_num]


# ---------------------------------------------------------------------- knowledge functions
class SciPyKnowledgeFunctionDefinition(KnowledgeFunctionDefinition):
    """product of independent 1-D scipy.stats densities, one per dimension of the supporting
    region; parameters live under "TruncatedNormalKnowledgeFunctionDefinition" (the reference
    uses that child for both subclasses, knowledge_function_definitions.py:20)"""

    def __init__(self, params, supporting_region):
        super().__init__(supporting_region,
                         params.AddChild("TruncatedNormalKnowledgeFunctionDefinition"))
        self.SetDefaultDistributionParams(self.supporting_region.definition)

    def SetDefaultDistributionParams(self, supporting_region):
        raise NotImplementedError

    def GetDistFunc(self, region_params, region_range, random_state=None):
        raise NotImplementedError

    def GetDistFuncOfRegion(self, region_desc):
        return self.GetDistFunc(self.params[region_desc],
                                self.supporting_region.definition[region_desc],
                                np.random.RandomState(1000))

    def CalculateIntegral(self, region_boundaries):
        total = 1.0
        support = self.supporting_region.definition
        for name, (lo, hi) in region_boundaries.items():
            dist = self.GetDistFunc(self.params[name], support[name])
            total *= float(dist.cdf(hi) - dist.cdf(lo))
        return total

    def Sample(self, sampling_region, random_state):
        """-> (values by dimension, density of the knowledge function at the values,
        density of the proposal the values were drawn from)"""
        values, density = {}, 1.0
        if sampling_region:
            for name, bounds in sampling_region.items():
                x = random_state.uniform(low=bounds[0], high=bounds[1])
                density *= float(self.GetDistFunc(self.params[name], bounds, random_state).pdf(x))
                values[name] = x
            return values, density, 1.0 / CalculateRegionBoundariesArea(sampling_region)
        for name, bounds in self.supporting_region.definition.items():
            dist = self.GetDistFunc(self.params[name], bounds, random_state)
            x = float(dist.rvs(size=1)[0])
            density *= float(dist.pdf(x))
            values[name] = x
        return values, density, density


class UniformKnowledgeFunctionDefinition(SciPyKnowledgeFunctionDefinition):
    def GetDistFunc(self, region_params, region_range, random_state=None):
        import scipy.stats
        dist = scipy.stats.uniform(loc=region_range[0], scale=region_range[1] - region_range[0])
        dist.random_state = random_state
        return dist

    def SetDefaultDistributionParams(self, supporting_region):
        for name in supporting_region:
            self.params[name] = None


class TruncatedNormalKnowledgeFunctionDefinition(SciPyKnowledgeFunctionDefinition):
    def GetDistFunc(self, region_params, region_range, random_state=None):
        import scipy.stats
        mean, std = region_params["Mean"], region_params["Std"]
        dist = scipy.stats.truncnorm(a=(region_range[0] - mean) / std,
                                     b=(region_range[1] - mean) / std, loc=mean, scale=std)
        dist.random_state = random_state
        return dist

    def SetDefaultDistributionParams(self, supporting_region):
        for name in supporting_region:
            self.params[name]["Mean", "Mean of truncated normal", 5]
            self.params[name]["Std", "Std of truncated normal", 1]


KNOWLEDGE_FUNCTIONS = {
    "TruncatedNormalKnowledgeFunctionDefinition": TruncatedNormalKnowledgeFunctionDefinition,
    "UniformKnowledgeFunctionDefinition": UniformKnowledgeFunctionDefinition,
}
RISK_FUNCTIONS = {"LinearKnowledgeFunctionDefinition": LinearKnowledgeFunctionDefinition}


# ---------------------------------------------------------------------- the behaviour space
class BehaviorSpace:
    def __init__(self, params):
        self._params = params.AddChild("BehaviorSpace")
        self._behavior_space_definition = self._params.AddChild("Definition")
        self._scenario_risk_function = None
        self._config_behavior_space()

    # -- definition (behavior_space.py:165-199)
    def _config_behavior_space(self):
        definition = self._behavior_space_definition
        self.model_type = definition["ModelType", "Model type over which behavior space is defined",
                                     "BehaviorIDMStochastic"]
        self._behavior_space_range_params = definition.AddChild("SpaceBoundaries")
        model_defaults = ParameterServer()
        self._model_from_model_type(self.model_type, model_defaults)

        def to_ranges(defaults, boundaries):
            for key, value in defaults.store.items():
                if _is_distribution(key):
                    boundaries[key, "Range for this distribution", list(DEFAULT_RANGE)]
                elif isinstance(value, ParameterServer):
                    to_ranges(value, boundaries[key])
                else:
                    boundaries[key, "Value", value]

        to_ranges(model_defaults, self._behavior_space_range_params)

        self._prior_knowledge_function_params = definition.AddChild("PriorKnowledgeFunction")
        name = self._prior_knowledge_function_params[
            "FunctionDefinition", "Class derived of KnowledgeFunctionDefinition that defines the "
            "prior knowledge", "TruncatedNormalKnowledgeFunctionDefinition"]
        ranges, _ = self._divide_distribution_ranges_and_fixed(self._behavior_space_range_params)
        self._prior_knowledge_region = PriorKnowledgeRegion(ranges)
        self._prior_knowledge_function_definition = KNOWLEDGE_FUNCTIONS[name](
            self._prior_knowledge_function_params, self._prior_knowledge_region)
        self._prior_knowledge_function = PriorKnowledgeFunction(
            self._prior_knowledge_region, self._prior_knowledge_function_definition,
            self._prior_knowledge_function_params)

    def _divide_distribution_ranges_and_fixed(self, behavior_space_range_params):
        """-> ({"Model::XDistribution": (lo, hi)}, nested dict of the fixed-value entries)"""
        ranges, fixed = {}, {}
        for path, val in _distribution_entries(behavior_space_range_params):
            if len(val) == 2:
                ranges[DELIMITER.join(path)] = tuple(val)
            else:
                level = fixed
                for part in path[:-1]:
                    level = level.setdefault(part, {})
                level[path[-1]] = val
        return ranges, fixed

    def _model_from_model_type(self, model_type, params):
        return MODEL_TYPES[model_type](params), params

    def get_prior_knowledge_function(self):
        return self._prior_knowledge_function

    def get_scenario_risk_function(self):
        if self._scenario_risk_function is None:
            fn_params = self._behavior_space_definition.AddChild("ScenarioRiskFunctionFunction")
            name = fn_params["FunctionName", "Name of scenario risk function definition",
                             "LinearKnowledgeFunctionDefinition"]
            template = RISK_FUNCTIONS[name](self._prior_knowledge_region, fn_params)
            self._scenario_risk_function = \
                self._prior_knowledge_function.CalculateScenarioRiskFunction(template)
        return self._scenario_risk_function

    # -- hypothesis sets (behavior_space.py:47-163)
    def get_default_hypothesis_parameters(self):
        hyp = self._params.AddChild("Hypothesis")
        hyp["RandomSeed", "Seed for hypothesis", 1000]
        model_type = hyp["HypothesisModel", "Model used as behavior model for hypothesis",
                         "BehaviorHypothesisIDM"]
        if len(hyp[model_type]) == 0:
            defaults = ParameterServer()
            self._model_from_model_type(model_type, defaults)
            hyp[model_type] = defaults[model_type]
        partitions = hyp.AddChild("Partitions")
        for path, _ in _distribution_entries(self._behavior_space_range_params):
            _node(partitions, path[:-1])[path[-1], "Number of partitions", 1]
        return hyp.clone()

    def create_hypothesis_set(self, hypothesis_parameters=None):
        hyp = hypothesis_parameters or self.get_default_hypothesis_parameters()
        seed = hyp["RandomSeed"]
        model_type = hyp["HypothesisModel"]
        paths, choices = [], []
        for path, partition_num in _distribution_entries(hyp.AddChild("Partitions")):
            param_range = _node(self._behavior_space_range_params, path[:-1])[path[-1]]
            paths.append(path)
            if len(param_range) == 2:
                choices.append([_partition_of(param_range, partition_num, i)
                                for i in range(partition_num)])
            else:
                choices.append([list(param_range)])

        hypothesis_set, hypothesis_set_params = [], []
        for combination in itertools.product(*choices):
            model_params = self._behavior_space_range_params.clone()
            for path, bounds in zip(paths, combination):
                dist = model_params.AddChild(DELIMITER.join(path), delete=True)
                if len(bounds) == 2:
                    dist["DistributionType"] = "UniformDistribution1D"
                    dist["RandomSeed"] = seed
                    dist["LowerBound"], dist["UpperBound"] = bounds
                else:
                    dist["DistributionType"] = "FixedValue"
                    dist["FixedValue"] = [bounds[0]]
            model_params[model_type] = hyp.AddChild(model_type).clone()
            params = ParameterServer(json=model_params.ConvertToDict(), log_if_default=True)
            hypothesis, _ = self._model_from_model_type(model_type, params)
            hypothesis_set.append(hypothesis)
            hypothesis_set_params.append(params)
        return hypothesis_set, hypothesis_set_params

    def create_hypothesis_set_fixed_split(self, split):
        hyp = self.get_default_hypothesis_parameters()
        partitions = hyp.AddChild("Partitions")
        for path, _ in list(_distribution_entries(partitions)):
            _node(partitions, path[:-1])[path[-1]] = split
        return self.create_hypothesis_set(hyp)

    def create_cover_hypothesis(self):
        """one hypothesis over the whole space"""
        return self.create_hypothesis_set_fixed_split(split=1)

    def create_multiple_hypothesis_sets(self, splits):
        return {split: self.create_hypothesis_set_fixed_split(split=split) for split in splits}

    # -- sampling of behaviour parameters (behavior_space.py:24-45, 240-343)
    def sample_behavior_parameters(self, random_state=None):
        self._sampling_parameters = self._params.AddChild("Sampling")
        seed = self._sampling_parameters["RandomSeed", "Seed for parameter sampling", 1000]
        self.random_state = random_state or np.random.RandomState(seed)

        partitioned, range_params = self._check_and_apply_partitioning(
            self._behavior_space_range_params, self._sampling_parameters)
        region = None
        if partitioned:
            region, _ = self._divide_distribution_ranges_and_fixed(range_params)
        else:
            range_params = self._behavior_space_range_params

        means, knowledge_probability, importance_sampling_probability = \
            self._sample_mean_params_from_prior_knowledge_function(range_params, region)
        sampled = self._sample_param_variations_from_param_means(means, range_params,
                                                                 self._sampling_parameters)
        return sampled, self.model_type, knowledge_probability, importance_sampling_probability

    def _check_and_apply_partitioning(self, behavior_space_range_params, sampling_params):
        partitioned_params = behavior_space_range_params.clone()
        partitioned = False
        for path, param_range in list(_distribution_entries(partitioned_params)):
            if len(param_range) < 2:
                continue
            sampling = _node(sampling_params, path)
            count = sampling["Partitions", "Into how many equal parameter range partitions is "
                             "this range partitioned", None]
            if count:
                which = sampling["SelectedPartition",
                                 "Which of these partitions is selected for sampling", 0]
                _node(partitioned_params, path[:-1])[path[-1]] = \
                    _partition_of(param_range, count, which)
                partitioned = True
        return partitioned, partitioned_params

    def _sample_mean_params_from_prior_knowledge_function(self, range_params, sample_region):
        means = range_params.clone()
        values, p_knowledge, p_sampling = \
            self._prior_knowledge_function_definition.Sample(sample_region, self.random_state)
        for name, value in values.items():
            path = name.split(DELIMITER)
            _node(means, path[:-1])[path[-1]] = value
        return means, p_knowledge, p_sampling

    def _sample_param_variations_from_param_means(self, mean_space_params, range_params,
                                                  sampling_params):
        out = ParameterServer(log_if_default=True)
        for key, mean in mean_space_params.store.items():
            child = sampling_params[key]
            param_range = range_params[key]
            if _is_distribution(key):
                kind = child["DistributionType", "Distribution type for sampling",
                             "UniformDistribution1D"]
                if "Fixed" in kind or len(param_range) == 1:
                    out[key] = self._get_fixed_dist_params(mean, param_range, child)
                elif "Uniform" in kind:
                    out[key] = self._sample_uniform_dist_params(mean, param_range, child)
                elif "Normal" in kind:
                    out[key] = self._sample_normal_dist_params(mean, param_range, child)
                else:
                    raise ValueError(f"{key}: unknown sampling DistributionType {kind!r}")
            elif isinstance(mean, ParameterServer):
                out[key] = self._sample_param_variations_from_param_means(mean, param_range, child)
            else:
                out[key] = self._sample_non_distribution_params(mean, param_range, child)
            if len(child) == 0:
                del sampling_params[key]
        return out

    def _sample_non_distribution_params(self, sampled_mean, value_range, sampling_params):
        if isinstance(value_range, list):
            if len(value_range) > 1:
                return self.random_state.uniform(value_range[0], value_range[1])
            return value_range[0]
        return value_range

    def _sample_uniform_dist_params(self, sampled_mean, value_range, sampling_params):
        """a uniform distribution of random width around the mean, shifted to stay inside the range"""
        lo_w, hi_w = sampling_params["Width", "What minimum and maximum width should sampled "
                                     "distribution have", [0.1, 0.3]]
        width = self.random_state.uniform(lo_w, hi_w)
        below = min(sampled_mean - value_range[0],
                    max(width / 2.0, width - (value_range[1] - sampled_mean)))
        out = ParameterServer(log_if_default=True)
        out["DistributionType"] = "UniformDistribution1D"
        out["LowerBound"] = sampled_mean - below
        out["UpperBound"] = sampled_mean + (width - below)
        out["RandomSeed"] = sampling_params["RandomSeed", "Seed for stochastic behavior", 1000]
        return out

    def _get_fixed_dist_params(self, sampled_mean, value_range, sampling_params):
        out = ParameterServer(log_if_default=True)
        out["DistributionType"] = "FixedValue"
        out["FixedValue"] = sampled_mean
        return out

    def _sample_normal_dist_params(self, sampled_mean, value_range, sampling_params):
        lo_s, hi_s = sampling_params["StdRange", "What minimum and maximum standard deviation "
                                     "should sampled distribution have", [0.1, 0.3]]
        out = ParameterServer(log_if_default=True)
        out["DistributionType"] = "NormalDistribution1D"
        out["Mean"] = sampled_mean
        out["StdDev"] = self.random_state.uniform(lo_s, hi_s)
        out["RandomSeed"] = sampling_params["RandomSeed", "Seed for stochastic behavior", 1000]
        return out
