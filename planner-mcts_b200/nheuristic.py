"""Neural warm start of the risk-constrained search: the Python side
of the seam in host/node_prior.hpp, with the reference's class names
(behavior_uct_nheuristic_risk_constraint.hpp:23-60, python_wrapper/python_planner_uct.cpp).

    observer = NearestObserver(params)
    converter = NNToValueConverterSequential(num_actions)
    planner = BehaviorUCTNHeuristicRiskConstraint(params, hypotheses, risk_function,
                                                  "model.pt", observer, converter, device=0)
    planner.Plan(0.2, observed_world)

The model (a TorchScript file, a torch.nn.Module or any callable observations -> outputs) maps
observations [n, observation_size] to [n, 3 * num_actions] (+ num_actions with a policy head).
The search calls it once per WAVE with the observations of every node the wave created -- one
batched forward pass on the GPU -- where the reference evaluates one node at a time on the CPU
(mcts_neural_cost_constrained_statistic.hpp:175-198).  bark-ml is not in the container: observer
and converter are restated from their parameters and call sites; parity is unpinned.
"""
import ctypes as C
import math

import numpy as np

from . import _abi
from .planner import (PLANNER_NHEURISTIC_RISK_CONSTRAINT, AgentState, BehaviorUCTRiskConstraint)

PRIOR_POLICY, PRIOR_VALUES = 1, 2
NODE_PRIOR_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int,
                            C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float),
                            C.POINTER(C.c_float))


def _get(params, path, default):
    if params is None:
        return default
    if hasattr(params, "AddChild"):
        return params[path, "", default]
    return params.get(path, default)


class NearestObserver:
    """ML::NearestObserver (tests/data/nheuristic_test.json:132-149): [x, y, theta, v] of the ego
    and of its NNearestAgents nearest agents within MaxDist, nearest first, each normalised to
    [0, 1] (x / y over the bounding box of the map's road polygon); zero rows for missing agents.
    The same function as host/node_prior.hpp NearestObserver (tests compare the two)."""

    def __init__(self, params=None):
        k = "ML::NearestObserver::"
        self.n_nearest = int(_get(params, k + "NNearestAgents", 4))
        self.min_vel = float(_get(params, k + "MinVel", 0.0))
        self.max_vel = float(_get(params, k + "MaxVel", 50.0))
        self.max_dist = float(_get(params, k + "MaxDist", 75.0))
        self.min_theta = float(_get(params, k + "MinTheta", -math.pi))
        self.max_theta = float(_get(params, k + "MaxTheta", math.pi))

    @property
    def observation_size(self):
        return (self.n_nearest + 1) * 4

    def _row(self, a, box):
        def norm(v, lo, hi):
            return np.float32((float(v) - lo) / (hi - lo)) if hi > lo else np.float32(0.0)
        return [norm(np.float32(a.x), box[0], box[1]), norm(np.float32(a.y), box[2], box[3]),
                norm(np.float32(a.theta), self.min_theta, self.max_theta),
                norm(np.float32(a.v), self.min_vel, self.max_vel)]

    def Observe(self, observed_world):
        m = observed_world.map
        xs = [float(v) for v in list(m.road_x)[:m.n_road_pts]]
        ys = [float(v) for v in list(m.road_y)[:m.n_road_pts]]
        box = (min(xs), max(xs), min(ys), max(ys))
        ego = next(a for a in observed_world.agents if a.id == observed_world.ego_id)
        out = np.zeros(self.observation_size, dtype=np.float32)
        out[:4] = self._row(ego, box)
        near = []
        for a in observed_world.agents:
            if a is ego:
                continue
            dx = float(np.float32(a.x)) - float(np.float32(ego.x))
            dy = float(np.float32(a.y)) - float(np.float32(ego.y))
            d = math.sqrt(dx * dx + dy * dy)
            if d < self.max_dist:
                near.append((d, a))
        near.sort(key=lambda t: t[0])
        for i, (_, a) in enumerate(near[:self.n_nearest]):
            out[(i + 1) * 4:(i + 2) * 4] = self._row(a, box)
        return out


class NNToValueConverterSequential:
    """model output [n, 3 * num_actions] = envelope risk | collision risk | return, per action
    (the order of the reference's training targets,
    tests/py_behavior_uct_nheuristic_risk_constraint_test.py:112-125); with_policy adds a fourth
    block of non-negative policy weights"""

    def __init__(self, num_actions, with_policy=False):
        self.num_actions = int(num_actions)
        self.with_policy = bool(with_policy)

    @property
    def output_size(self):
        return self.num_actions * (4 if self.with_policy else 3)

    def ConvertToValueMap(self, nn_output):
        out = np.asarray(nn_output, dtype=np.float32).reshape(-1, self.output_size)
        n = self.num_actions
        values = {"EnvelopeRisk": out[:, :n], "CollisionRisk": out[:, n:2 * n],
                  "Return": out[:, 2 * n:3 * n]}
        if self.with_policy:
            values["Policy"] = np.maximum(out[:, 3 * n:4 * n], 0.0)
        return values


class TorchNodePrior:
    """a torch model evaluated on `device` for a whole batch of observations"""

    def __init__(self, model, converter, device="cuda:0"):
        import torch
        self.torch = torch
        self.device = torch.device(device)
        if isinstance(model, str):
            model = torch.jit.load(model, map_location=self.device)
        self.model = model.to(self.device).eval()
        self.converter = converter
        self.batches = 0
        self.rows = 0

    def __call__(self, observations):
        torch = self.torch
        with torch.no_grad():
            x = torch.from_numpy(observations)
            if self.device.type == "cuda":
                x = x.pin_memory().to(self.device, non_blocking=True)
            y = self.model(x).float().cpu().numpy()
        self.batches += 1
        self.rows += len(observations)
        return self.converter.ConvertToValueMap(y)


class BehaviorUCTNHeuristicRiskConstraint(BehaviorUCTRiskConstraint):
    """BehaviorUCTNHeuristicRiskConstraint(params, hypotheses, scenario_risk_function,
    model_file_name, observer, nn_to_value_converter).  `model_file_name` may also be a
    torch.nn.Module or a callable observations[n, d] -> {"Return", "EnvelopeRisk",
    "CollisionRisk"[, "Policy"]: [n, num_actions]}."""
    KIND = PLANNER_NHEURISTIC_RISK_CONSTRAINT

    def __init__(self, params=None, hypotheses=(), scenario_risk_function=None,
                 model_file_name=None, observer=None, nn_to_value_converter=None, device=0,
                 _lib=None, _create=None):
        super().__init__(params, hypotheses, scenario_risk_function, device, _lib, _create)
        self.model_file_name = model_file_name
        self.observer = observer or NearestObserver(self.params)
        self.nn_to_value_converter = nn_to_value_converter
        self._device = device
        self._prior = None
        self._prior_error = None
        self._lib.bmcts_planner_set_node_prior.argtypes = [C.c_void_p, NODE_PRIOR_FN, C.c_void_p]
        self._lib.bmcts_planner_observation_size.argtypes = [C.c_void_p]
        self._lib.bmcts_planner_observe.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(AgentState),
                                                    C.c_int, C.POINTER(C.c_float), C.c_int]
        self._lib.bmcts_planner_node_prior_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64),
                                                             C.POINTER(C.c_uint64)]
        self._callback = NODE_PRIOR_FN(self._evaluate)   # must outlive the native planner
        self._check(self._lib.bmcts_planner_set_node_prior(self._h, self._callback, None))

    # reference getters (behavior_uct_nheuristic_risk_constraint.hpp:44-46)
    def GetObserver(self):
        return self.observer

    def GetModelFileName(self):
        return self.model_file_name

    def GetNNToValueConverter(self):
        return self.nn_to_value_converter

    observer_ = property(lambda self: self.observer)
    model_filename = property(lambda self: self.model_file_name)

    def _make_prior(self):
        m = self.model_file_name
        if m is None:
            raise ValueError("BehaviorUCTNHeuristicRiskConstraint needs a model")
        is_torch = isinstance(m, str)
        if not is_torch:
            try:
                import torch
                is_torch = isinstance(m, torch.nn.Module)
            except ImportError:
                pass
        if not is_torch:
            return m  # plain callable
        if self.nn_to_value_converter is None:
            raise ValueError("a torch model needs an nn_to_value_converter")
        import torch
        dev = f"cuda:{self._device}" if torch.cuda.is_available() else "cpu"
        return TorchNodePrior(m, self.nn_to_value_converter, dev)

    def _evaluate(self, _user, n, obs_dim, obs, n_actions, policy, ret, env, col):
        try:
            if self._prior is None:
                self._prior = self._make_prior()
            x = np.ctypeslib.as_array(obs, shape=(n, obs_dim)).copy()
            values = self._prior(x)
            for key, v in values.items():
                if np.shape(v) != (n, n_actions):
                    raise ValueError(f"node prior {key!r} has shape {np.shape(v)}, the search "
                                     f"expects ({n}, {n_actions}) = (nodes, ego actions of this world)")
            mask = 0
            if "Policy" in values:
                np.ctypeslib.as_array(policy, shape=(n, n_actions))[:] = values["Policy"]
                mask |= PRIOR_POLICY
            if all(k in values for k in ("Return", "EnvelopeRisk", "CollisionRisk")):
                np.ctypeslib.as_array(ret, shape=(n, n_actions))[:] = values["Return"]
                np.ctypeslib.as_array(env, shape=(n, n_actions))[:] = values["EnvelopeRisk"]
                np.ctypeslib.as_array(col, shape=(n, n_actions))[:] = values["CollisionRisk"]
                mask |= PRIOR_VALUES
            return mask
        except Exception as e:  # no exceptions across the C ABI
            self._prior_error = e
            return -1

    def Plan(self, delta_time, observed_world):
        self._prior_error = None
        try:
            return super().Plan(delta_time, observed_world)
        except _abi.BmctsError:
            if self._prior_error is not None:
                raise self._prior_error
            raise

    @property
    def observation_size(self):
        return int(self._lib.bmcts_planner_observation_size(self._h))

    def observe(self, observed_world):
        """the observation the search builds for the root of this world (native observer)"""
        self._check(self._lib.bmcts_planner_set_map(self._h, C.byref(observed_world.map)))
        self._map = observed_world.map
        n = len(observed_world.agents)
        arr = (AgentState * n)()
        for i, a in enumerate(observed_world.agents):
            arr[i] = AgentState(a.id, a.x, a.y, a.theta, a.v, a.length, a.width, a.rear, a.last_acc)
        out = (C.c_float * self.observation_size)()
        got = self._lib.bmcts_planner_observe(self._h, int(observed_world.ego_id), arr, n, out,
                                              len(out))
        if got < 0:
            raise _abi.BmctsError(got, self._lib.bmcts_planner_last_error(self._h).decode())
        return np.array(out[:got], dtype=np.float32)

    @property
    def node_prior_stats(self):
        """(calls of the model, nodes evaluated) since construction"""
        calls, nodes = C.c_uint64(), C.c_uint64()
        self._lib.bmcts_planner_node_prior_stats(self._h, C.byref(calls), C.byref(nodes))
        return int(calls.value), int(nodes.value)

    def __getstate__(self):
        st = super().__getstate__()
        st.update(model_file_name=self.model_file_name, observer=self.observer,
                  nn_to_value_converter=self.nn_to_value_converter, device=self._device)
        return st

    def __setstate__(self, st):
        self.__init__(st["params"], st["hypotheses"], st["scenario_risk_function"],
                      st["model_file_name"], st["observer"], st["nn_to_value_converter"],
                      st["device"])
        if st["tree_index"]:
            self.set_tree_index(st["tree_index"])
