"""Python binding of the BehaviorUCT* planners (ctypes over include/bmcts_planner_c.h).

Mirrors the reference's pybind classes (python_wrapper/python_planner_uct.cpp:131-248):
`BehaviorUCTHypothesis(params, hypotheses)`, `BehaviorUCTRiskConstraint(params, hypotheses)`,
`BehaviorUCTCooperative(params)`; `Plan(delta_time, observed_world)`; `last_*` debug infos.
`params` is a dict of BARK ParamServer keys ("BehaviorUctBase::Mcts::MaxNumIterations": 500);
`observed_world` is an `ObservedWorld` (agents + map geometry).
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import _abi

PLANNER_HYPOTHESIS, PLANNER_RISK_CONSTRAINT, PLANNER_COOPERATIVE = 0, 1, 2
PLANNER_NHEURISTIC_RISK_CONSTRAINT = 3
PLAN_MAX_TRAJ = 64


class AgentState(C.Structure):
    _fields_ = [("id", C.c_uint32), ("x", C.c_float), ("y", C.c_float), ("theta", C.c_float),
                ("v", C.c_float), ("length", C.c_float), ("width", C.c_float), ("rear", C.c_float),
                ("last_acc", C.c_float)]


class EdgeInfo(C.Structure):
    _fields_ = [("agent_id", C.c_uint32), ("depth", C.c_uint32), ("action", C.c_uint32),
                ("weight", C.c_float)]


class PlanResult(C.Structure):
    _fields_ = [
        ("best_action", C.c_uint32), ("num_iterations", C.c_uint32), ("num_nodes", C.c_uint32),
        ("num_actions", C.c_uint32), ("solution_time_ms", C.c_double),
        ("return_values", C.c_double * _abi.MAX_EGO_ACTIONS),
        ("visit_counts", C.c_double * _abi.MAX_EGO_ACTIONS),
        ("cost_values", (C.c_double * _abi.MAX_EGO_ACTIONS) * 2),
        ("policy", C.c_double * _abi.MAX_EGO_ACTIONS),
        ("expected_risk", C.c_double * 2), ("scenario_risk", C.c_double * 2),
        ("action_acc", C.c_float), ("action_steering", C.c_float),
        ("max_tree_depth", C.c_int32), ("root_expanded_actions", C.c_int32 * _abi.MAX_AGENTS),
        ("n_traj", C.c_int32), ("trajectory", C.c_double * (PLAN_MAX_TRAJ * 5)),
        ("engine_launches", C.c_uint64),
    ]


PLANNER_SYMBOLS = [
    "bmcts_planner_create", "bmcts_planner_destroy", "bmcts_planner_set_param",
    "bmcts_planner_add_hypothesis", "bmcts_planner_set_scenario_risk", "bmcts_planner_set_map", "bmcts_planner_set_tree_index",
    "bmcts_planner_plan", "bmcts_planner_get_beliefs", "bmcts_planner_root_stats",
    "bmcts_planner_decide_from_root_stats", "bmcts_planner_parameter_sampling_probability",
    "bmcts_planner_last_error", "bmcts_planner_set_node_prior", "bmcts_planner_observation_size",
    "bmcts_planner_observe", "bmcts_planner_node_prior_stats", "bmcts_planner_update_beliefs",
    "bmcts_planner_reset_beliefs", "bmcts_planner_allreduce_root_stats",
    "bmcts_planner_set_true_behavior", "bmcts_planner_last_edges", "bmcts_planner_last_states",
    "bmcts_planner_set_agent_goal",
]


def declare_planner_api(lib):
    P = C.POINTER
    lib.bmcts_planner_create.argtypes = [C.c_int, C.c_int, P(C.c_void_p)]
    lib.bmcts_planner_destroy.argtypes = [C.c_void_p]
    lib.bmcts_planner_destroy.restype = None
    lib.bmcts_planner_set_param.argtypes = [C.c_void_p, C.c_char_p, P(C.c_double), C.c_int]
    lib.bmcts_planner_add_hypothesis.argtypes = [C.c_void_p, P(_abi.Hypothesis), C.c_int, C.c_int,
                                                 C.c_float, C.c_float]
    lib.bmcts_planner_set_scenario_risk.argtypes = [C.c_void_p, P(C.c_double), C.c_int]
    lib.bmcts_planner_set_map.argtypes = [C.c_void_p, P(_abi.Scenario)]
    lib.bmcts_planner_set_tree_index.argtypes = [C.c_void_p, C.c_uint32]
    lib.bmcts_planner_plan.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_uint32,
                                       P(AgentState), C.c_int, P(PlanResult)]
    lib.bmcts_planner_get_beliefs.argtypes = [C.c_void_p, C.c_uint32, P(C.c_double), C.c_int]
    lib.bmcts_planner_root_stats.argtypes = [C.c_void_p, P(C.c_double), C.c_int]
    lib.bmcts_planner_decide_from_root_stats.argtypes = [C.c_void_p, P(C.c_double), C.c_int,
                                                         P(PlanResult)]
    lib.bmcts_planner_update_beliefs.argtypes = [C.c_void_p, C.c_double, C.c_uint32, P(AgentState),
                                                 C.c_int]
    lib.bmcts_planner_reset_beliefs.argtypes = [C.c_void_p]
    lib.bmcts_planner_parameter_sampling_probability.argtypes = [C.c_void_p, C.c_int]
    lib.bmcts_planner_parameter_sampling_probability.restype = C.c_double
    lib.bmcts_planner_last_error.argtypes = [C.c_void_p]
    lib.bmcts_planner_last_error.restype = C.c_char_p
    lib.bmcts_planner_set_true_behavior.argtypes = [C.c_void_p, C.c_uint32, P(_abi.Hypothesis)]
    lib.bmcts_planner_set_agent_goal.argtypes = [C.c_void_p, C.c_uint32, P(_abi.Goal)]
    lib.bmcts_planner_last_edges.argtypes = [C.c_void_p, P(EdgeInfo), C.c_int]
    lib.bmcts_planner_last_states.argtypes = [C.c_void_p, P(C.c_uint32), P(C.c_uint32), P(C.c_float),
                                              C.c_int, P(C.c_int)]
    if hasattr(lib, "bmcts_planner_allreduce_root_stats"):   # libbmcts.so only (needs a GPU)
        lib.bmcts_planner_allreduce_root_stats.argtypes = [C.c_void_p, C.c_void_p, P(PlanResult)]
    return lib


@dataclass
class Agent:
    id: int
    x: float
    y: float
    theta: float = 0.0
    v: float = 0.0
    length: float = 4.0
    width: float = 2.0
    rear: float = 1.25
    last_acc: float = 0.0
    # the agent's own behaviour model (a HypothesisIDM / _abi.Hypothesis region), if known: the
    # omniscient planners (UseTrueBehaviorsAsHypothesis) predict the agent with it
    behavior_model: object = None
    # the agent's own goal definition (an _abi.Goal), if it has one: the cooperative planner
    # evaluates every agent against its own goal
    goal: object = None


@dataclass
class ObservedWorld:
    ego_id: int
    agents: List[Agent]
    map: object  # _abi.Scenario holding corridors, road polygon and goal
    world_time: float = 0.0


@dataclass
class HypothesisIDM:
    """BehaviorHypothesisIDM: parameter region + likelihood histogram settings
    (hypothesis/idm/hypothesis_idm.cpp:27-38)."""
    region: object  # _abi.Hypothesis
    num_samples: int = 10000
    num_buckets: int = 100
    buckets_lower_bound: float = -10.0
    buckets_upper_bound: float = 10.0

    @property
    def parameter_regions(self):
        """BehaviorHypothesisIDM::GetParameterRegions: {distribution name: (lower, upper)}"""
        names = ["BehaviorIDMStochastic::HeadwayDistribution", "BehaviorIDMStochastic::SpacingDistribution",
                 "BehaviorIDMStochastic::MaxAccDistribution", "BehaviorIDMStochastic::DesiredVelDistribution",
                 "BehaviorIDMStochastic::ComftBrakingDistribution",
                 "BehaviorIDMStochastic::CoolnessFactorDistribution"]
        return {n: (float(self.region.lo[i]), float(self.region.hi[i])) for i, n in enumerate(names)}

    # pickling (python_wrapper/python_planner_uct.cpp:206-248 pickles hypotheses with their
    # parameters): the ctypes region travels as its raw bytes
    def __getstate__(self):
        d = dict(self.__dict__)
        d["region"] = bytes(self.region)
        return d

    def __setstate__(self, d):
        d = dict(d)
        d["region"] = _abi.Hypothesis.from_buffer_copy(d["region"])
        self.__dict__.update(d)


class _BehaviorUCT:
    KIND = PLANNER_HYPOTHESIS

    def __init__(self, params=None, hypotheses=(), device=0, _lib=None, _create=None):
        self._lib = declare_planner_api(_lib or _abi.load_library())
        self._h = C.c_void_p()
        if _create is not None:
            st = _create(self.KIND, C.byref(self._h))
        else:
            st = self._lib.bmcts_planner_create(self.KIND, int(device), C.byref(self._h))
        if st != _abi.OK:
            raise _abi.BmctsError(st, "bmcts_planner_create")
        # a flat {"A::B::key": value} dict or a parameters.ParameterServer tree
        self.params = params if hasattr(params, "flatten") else dict(params or {})
        for key, val in self.params.items():
            if val is None or isinstance(val, str):
                continue  # names of model / function types: not planner parameters
            vals = np.atleast_1d(np.asarray(val, dtype=np.float64)).ravel()
            arr = (C.c_double * len(vals))(*vals)
            self._check(self._lib.bmcts_planner_set_param(self._h, key.encode(), arr, len(vals)))
        self.hypotheses = list(hypotheses)
        for h in self.hypotheses:
            self._check(self._lib.bmcts_planner_add_hypothesis(
                self._h, C.byref(h.region), h.num_samples, h.num_buckets,
                h.buckets_lower_bound, h.buckets_upper_bound))
        self._map = None
        self.last = None

    def _check(self, st):
        if st != _abi.OK:
            raise _abi.BmctsError(st, self._lib.bmcts_planner_last_error(self._h).decode())

    def __del__(self):
        try:
            if self._h.value:
                self._lib.bmcts_planner_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def set_tree_index(self, i):
        self._tree_index = int(i)
        self._check(self._lib.bmcts_planner_set_tree_index(self._h, int(i)))

    # pickling / deepcopy: parameters + hypotheses (+ risk function) are kept, the native
    # planner is rebuilt; beliefs and debug infos of the last Plan() are not carried over
    # (the reference does not pickle its belief tracker either, python_planner_uct.cpp:145-166)
    def __getstate__(self):
        return {"params": self.params, "hypotheses": self.hypotheses,
                "tree_index": getattr(self, "_tree_index", 0),
                "scenario_risk_function": getattr(self, "scenario_risk_function", None)}

    def __setstate__(self, st):
        if isinstance(self, BehaviorUCTCooperative):
            self.__init__(st["params"])
        elif isinstance(self, BehaviorUCTRiskConstraint):
            self.__init__(st["params"], st["hypotheses"], st["scenario_risk_function"])
        else:
            self.__init__(st["params"], st["hypotheses"])
        if st["tree_index"]:
            self.set_tree_index(st["tree_index"])

    def Plan(self, delta_time, observed_world):
        """Plan(min_planning_time, observed_world) -> trajectory ndarray [n, 5] = [t,x,y,theta,v]"""
        if self._map is not observed_world.map:
            self._check(self._lib.bmcts_planner_set_map(self._h, C.byref(observed_world.map)))
            self._map = observed_world.map
        n = len(observed_world.agents)
        arr = (AgentState * n)()
        for i, a in enumerate(observed_world.agents):
            arr[i] = AgentState(a.id, a.x, a.y, a.theta, a.v, a.length, a.width, a.rear, a.last_acc)
            if a.goal is not None:
                self._check(self._lib.bmcts_planner_set_agent_goal(self._h, int(a.id), C.byref(a.goal)))
            if a.behavior_model is not None:
                region = getattr(a.behavior_model, "region", a.behavior_model)
                self._check(self._lib.bmcts_planner_set_true_behavior(self._h, int(a.id), C.byref(region)))
        res = PlanResult()
        self._check(self._lib.bmcts_planner_plan(self._h, float(delta_time),
                                                 float(observed_world.world_time),
                                                 int(observed_world.ego_id), arr, n, C.byref(res)))
        self.last = res
        return self.last_trajectory

    # ---- debug infos (python_planner_uct.cpp:145-166, 206-248)
    def _vec(self, field_, n=None):
        n = self.last.num_actions if n is None else n
        return np.array(list(getattr(self.last, field_))[:n], dtype=np.float64)

    @property
    def last_trajectory(self):
        return np.array(self.last.trajectory[:self.last.n_traj * 5]).reshape(-1, 5)

    @property
    def last_action(self):
        return np.array([self.last.action_acc, self.last.action_steering])

    @property
    def last_macro_action(self):
        return int(self.last.best_action)

    def GetLastMacroAction(self):
        return self.last_macro_action

    @property
    def last_return_values(self):
        v = self._vec("return_values")
        return {a: float(x) for a, x in enumerate(v) if not np.isnan(x)}

    @property
    def last_visit_counts(self):
        return self._vec("visit_counts")

    @property
    def last_num_iterations(self):
        return int(self.last.num_iterations)

    @property
    def last_num_nodes(self):
        return int(self.last.num_nodes)

    @property
    def last_solution_time(self):
        return float(self.last.solution_time_ms)

    @property
    def last_max_tree_depth(self):
        return int(self.last.max_tree_depth)

    @property
    def last_root_expanded_actions(self):
        return list(self.last.root_expanded_actions)

    @property
    def engine_launches(self):
        return int(self.last.engine_launches)

    @property
    def last_extracted_mcts_edges(self):
        """python_planner_uct.cpp:175,202,265: [(agent id, depth, action, action weight), ...] of
        the last search tree down to MaxExtractionDepth (ExtractEdgeInfo, default true)"""
        n = self._lib.bmcts_planner_last_edges(self._h, None, 0)
        buf = (EdgeInfo * max(1, n))()
        self._lib.bmcts_planner_last_edges(self._h, buf, n)
        return [(int(e.agent_id), int(e.depth), int(e.action), float(e.weight)) for e in buf[:n]]

    @property
    def last_extracted_mcts_states(self):
        """[(depth, {agent id: (x, y, theta, v)}), ...] of the extracted nodes (ExtractStateInfo)"""
        A = C.c_int(0)
        n = self._lib.bmcts_planner_last_states(self._h, None, None, None, 0, C.byref(A))
        if n == 0:
            return []
        depth = (C.c_uint32 * n)()
        ids = (C.c_uint32 * (n * A.value))()
        pose = (C.c_float * (n * A.value * 4))()
        self._lib.bmcts_planner_last_states(self._h, depth, ids, pose, n, C.byref(A))
        out = []
        for i in range(n):
            agents = {}
            for a in range(A.value):
                aid = ids[i * A.value + a]
                if aid != 0xFFFFFFFF:
                    agents[int(aid)] = tuple(pose[(i * A.value + a) * 4:(i * A.value + a) * 4 + 4])
            out.append((int(depth[i]), agents))
        return out

    @property
    def ego_behavior(self):
        """python_planner_uct.cpp (ego_behavior): the ego's macro-action model as the planner builds
        it from BehaviorUctBase::EgoBehavior::* -- the motion primitives in action-index order"""
        b = "BehaviorUctBase::EgoBehavior::"
        acc = self.params.get(b + "AccelerationInputs", [0.0, 1.0, 4.0, -1.0, -8.0])
        prims = [("PrimitiveConstAccStayLane", float(a)) for a in acc]
        if self.params.get(b + "AddLaneChangeActions", 1):
            prims += [("PrimitiveConstAccChangeToLeft", 0.0), ("PrimitiveConstAccChangeToRight", 0.0)]
        return {"model": "BehaviorMPMacroActions", "motion_primitives": prims,
                "valid_primitives_at_last_plan": int(self.last.num_actions)}

    def root_stats(self):
        buf = (C.c_double * 128)()
        n = self._lib.bmcts_planner_root_stats(self._h, buf, 128)
        return np.array(buf[:n], dtype=np.float64)

    def decide_from_root_stats(self, stats):
        stats = np.ascontiguousarray(stats, dtype=np.float64)
        res = PlanResult()
        self._check(self._lib.bmcts_planner_decide_from_root_stats(
            self._h, stats.ctypes.data_as(C.POINTER(C.c_double)), len(stats), C.byref(res)))
        self.last = res
        return int(res.best_action)


    def allreduce_root_stats(self, comm):
        """sum the root statistics of the last Plan() over the ranks of `comm`
        (root_parallel.RootComm: one ncclAllReduce behind the C ABI) and decide from the merged
        statistics; returns the action every rank agrees on"""
        res = PlanResult()
        self._check(self._lib.bmcts_planner_allreduce_root_stats(self._h, comm.handle, C.byref(res)))
        self.last = res
        return int(res.best_action)


class BehaviorUCTHypothesis(_BehaviorUCT):
    KIND = PLANNER_HYPOTHESIS

    def current_beliefs(self, agent_id):
        buf = (C.c_double * _abi.MAX_HYPOTHESES)()
        n = self._lib.bmcts_planner_get_beliefs(self._h, int(agent_id), buf, _abi.MAX_HYPOTHESES)
        return np.array(buf[:n])

    def parameter_sampling_probability(self, h):
        return float(self._lib.bmcts_planner_parameter_sampling_probability(self._h, int(h)))


class BeliefCalculator:
    """BeliefCalculator(params, hypotheses) (python_planner_uct.cpp:355-372,
    belief_calculator/belief_calculator.hpp:23-47): beliefs over the hypothesis set from a sequence
    of observed worlds, without planning.  The likelihoods are computed by the same engine call as
    inside the planners."""

    def __init__(self, params=None, hypotheses=(), device=0, _lib=None, _create=None):
        self._planner = BehaviorUCTHypothesis(params, hypotheses, device=device, _lib=_lib,
                                              _create=_create)
        self._agent_ids = []

    def BeliefUpdate(self, observed_world):
        pl = self._planner
        if pl._map is not observed_world.map:
            pl._check(pl._lib.bmcts_planner_set_map(pl._h, C.byref(observed_world.map)))
            pl._map = observed_world.map
        n = len(observed_world.agents)
        arr = (AgentState * n)()
        for i, a in enumerate(observed_world.agents):
            arr[i] = AgentState(a.id, a.x, a.y, a.theta, a.v, a.length, a.width, a.rear, a.last_acc)
        pl._check(pl._lib.bmcts_planner_update_beliefs(pl._h, float(observed_world.world_time),
                                                       int(observed_world.ego_id), arr, n))
        for a in observed_world.agents:
            if a.id != observed_world.ego_id and a.id not in self._agent_ids:
                self._agent_ids.append(a.id)

    def GetBeliefs(self):
        """{agent id: [belief per hypothesis]} for every other agent observed so far"""
        out = {}
        for agent_id in self._agent_ids:
            b = self._planner.current_beliefs(agent_id)
            if len(b):
                out[agent_id] = b.tolist()
        return out

    def GetBehaviorHypotheses(self):
        return list(self._planner.hypotheses)

    def GetParams(self):
        return self._planner.params

    def Reset(self):
        pl = self._planner
        pl._check(pl._lib.bmcts_planner_reset_beliefs(pl._h))
        self._agent_ids = []

    def __getstate__(self):
        return {"params": self._planner.params, "hypotheses": self._planner.hypotheses}

    def __setstate__(self, st):
        self.__init__(st["params"], st["hypotheses"])


class BehaviorUCTRiskConstraint(BehaviorUCTHypothesis):
    """BehaviorUCTRiskConstraint(params, hypotheses, scenario_risk_function)
    (behavior_uct_risk_constraint.hpp:56-65); scenario_risk_function is a
    risk_calculation.ScenarioRiskFunction or None"""
    KIND = PLANNER_RISK_CONSTRAINT

    def __init__(self, params=None, hypotheses=(), scenario_risk_function=None, device=0, _lib=None,
                 _create=None):
        super().__init__(params, hypotheses, device, _lib, _create)
        self.scenario_risk_function = scenario_risk_function
        if scenario_risk_function is not None:
            from .risk_calculation import HypothesisRegion
            dims = list(scenario_risk_function.risk_function_definition.GetSupportingRegion()
                        .GetDefinition())
            vals = [scenario_risk_function.CalculateIntegralValue(HypothesisRegion(h.region, dims))
                    for h in self.hypotheses]
            arr = (C.c_double * len(vals))(*vals)
            self._check(self._lib.bmcts_planner_set_scenario_risk(self._h, arr, len(vals)))

    @property
    def last_cost_values(self):
        names = ("envelope", "collision")
        out = {}
        for i, nm in enumerate(names):
            v = np.array(list(self.last.cost_values[i])[:self.last.num_actions])
            d = {a: float(x) for a, x in enumerate(v) if not np.isnan(x)}
            if d:
                out[nm] = d
        return out

    @property
    def last_expected_risk(self):
        return list(self.last.expected_risk)

    @property
    def last_scenario_risk(self):
        return list(self.last.scenario_risk)

    @property
    def last_policy_sampled(self):
        pol = {a: float(p) for a, p in enumerate(self._vec("policy")) if p > 0}
        return int(self.last.best_action), pol


class BehaviorUCTCooperative(_BehaviorUCT):
    KIND = PLANNER_COOPERATIVE

    def __init__(self, params=None, device=0, _lib=None, _create=None):
        super().__init__(params, (), device, _lib, _create)
