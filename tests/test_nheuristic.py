"""Neural warm start: observer, value converter, the batched node-prior
seam of the search and its effect on the ego statistic.  The reference's own test needs a trained
bark-ml model (tests/py_behavior_uct_nheuristic_risk_constraint_test.py); here the model is a
function with known outputs, so every effect can be asserted."""
import ctypes as C

import numpy as np
import pytest

import oracle
import planner_worlds as W
from conftest import has_gpu
from planner_mcts_b200 import _abi, planner as P
from planner_mcts_b200.nheuristic import (BehaviorUCTNHeuristicRiskConstraint, NearestObserver,
                                          NNToValueConverterSequential, TorchNodePrior)

ACTIONS = (0.0, 1.0, -1.0, -2.0)


def _params(iterations, wave=0, **extra):
    p = W.base_params(iterations, acc_inputs=ACTIONS)
    p.update({"BehaviorUctRiskConstraint::DefaultAvailableRisk": [0.1, 0.0],
              W.P0 + "Mcts::State::SplitSafeDistCollision": 1,
              W.P0 + "Mcts::State::EvaluationParameters::AddSafeDist": 1,
              W.P0 + "Mcts::Engine::WaveSize": wave})
    p.update(extra)
    return p


def _cpu(cls, params, *args, **kw):
    lib = oracle.planner_lib()
    return cls(params, W.hypotheses(500), None, *args, _lib=lib,
               _create=lambda kind, out: lib.oracle_planner_create(kind, out), **kw)


def _world():
    return W.observed_world(2, 10.0, 8.0, 2.0, goal_center=(150.0, -1.75))


class _Prior:
    """a model with known outputs that records the batches it was asked for"""

    def __init__(self, n_actions=4, policy=None, values=None):
        self.n_actions, self.policy, self.values = n_actions, policy, values
        self.batches = []

    def __call__(self, obs):
        self.batches.append(obs.copy())
        n, out = len(obs), {}
        if self.policy is not None:
            out["Policy"] = np.tile(np.asarray(self.policy, np.float32), (n, 1))
        if self.values is not None:
            for k, v in self.values.items():
                out[k] = np.tile(np.asarray(v, np.float32), (n, 1))
        return out


def test_observer_known_answer_and_native_twin():
    world = _world()
    pl = _cpu(BehaviorUCTNHeuristicRiskConstraint, _params(8), _Prior())
    obs = NearestObserver(pl.params)
    assert obs.observation_size == pl.observation_size == 20
    want = obs.Observe(world)
    got = pl.observe(world)
    assert got.dtype == np.float32 and np.array_equal(got, want)
    m = world.map
    xs, ys = list(m.road_x)[:m.n_road_pts], list(m.road_y)[:m.n_road_pts]
    ego = world.agents[0]
    assert got[0] == pytest.approx((ego.x - min(xs)) / (max(xs) - min(xs)), abs=1e-6)
    assert got[1] == pytest.approx((ego.y - min(ys)) / (max(ys) - min(ys)), abs=1e-6)
    assert got[2] == pytest.approx(0.5, abs=1e-6)                # theta 0 in [-pi, pi]
    assert got[3] == pytest.approx(ego.v / 50.0, abs=1e-6)
    # two other agents, nearest first; rows 3 and 4 are padding
    d = [np.hypot(a.x - ego.x, a.y - ego.y) for a in world.agents[1:]]
    first = world.agents[1 + int(np.argmin(d))]
    assert got[4] == pytest.approx((first.x - min(xs)) / (max(xs) - min(xs)), abs=1e-6)
    assert np.all(got[12:] == 0.0)
    # agents beyond MaxDist are not observed
    near = NearestObserver({"ML::NearestObserver::MaxDist": min(d) + 0.1})
    assert np.count_nonzero(near.Observe(world)[8:]) == 0


def test_value_converter_layout():
    conv = NNToValueConverterSequential(3, with_policy=True)
    assert conv.output_size == 12
    v = conv.ConvertToValueMap(np.arange(24, dtype=np.float32).reshape(2, 12) - 10.0)
    assert v["EnvelopeRisk"][0].tolist() == [-10.0, -9.0, -8.0]
    assert v["CollisionRisk"][1].tolist() == [5.0, 6.0, 7.0]
    assert v["Return"][0].tolist() == [-4.0, -3.0, -2.0]
    assert v["Policy"][0].tolist() == [0.0, 0.0, 1.0] and v["Policy"][1].tolist() == [11.0, 12.0, 13.0]
    assert set(NNToValueConverterSequential(3).ConvertToValueMap(np.zeros((1, 9)))) == \
        {"EnvelopeRisk", "CollisionRisk", "Return"}


def test_prior_is_asked_once_per_wave_for_all_new_nodes():
    prior = _Prior(policy=[1, 1, 1, 1])
    pl = _cpu(BehaviorUCTNHeuristicRiskConstraint, _params(64, wave=16), prior)
    world = _world()
    pl.Plan(0.2, world)
    calls, nodes = pl.node_prior_stats
    assert calls == len(prior.batches) == 1 + 64 // 16          # the root, then one per wave
    assert len(prior.batches[0]) == 1
    assert np.array_equal(prior.batches[0][0], NearestObserver(pl.params).Observe(world))
    assert all(1 <= len(b) <= 16 for b in prior.batches[1:])
    assert nodes == sum(len(b) for b in prior.batches) <= pl.last_num_nodes
    assert all(b.shape[1] == 20 and b.dtype == np.float32 for b in prior.batches)
    # observations of expanded nodes are worlds half a second later: the ego has moved
    assert not np.array_equal(prior.batches[1][0][:4], prior.batches[0][0][:4])


@pytest.mark.parametrize("action", [0, 1, 2, 3])
def test_exploration_policy_picks_the_first_action(action):
    policy = [0.0] * 4
    policy[action] = 1.0
    pl = _cpu(BehaviorUCTNHeuristicRiskConstraint, _params(1, wave=1), _Prior(policy=policy))
    pl.Plan(0.2, _world())
    counts = pl.last_visit_counts
    assert counts[action] == 1 and counts.sum() == 1


def test_value_prior_initialises_the_action_statistics():
    """PriorVisits virtual visits at the prior's values: with a heavy prior the root's action
    values stay at the prior, and the best action is the prior's best"""
    values = {"Return": [-500.0, -500.0, 50.0, -500.0], "EnvelopeRisk": [0.0] * 4,
              "CollisionRisk": [0.0] * 4}
    heavy = _params(32, wave=8, **{W.P0 + "Mcts::NeuralStatistic::PriorVisits": 100000})
    pl = _cpu(BehaviorUCTNHeuristicRiskConstraint, heavy, _Prior(values=values))
    pl.Plan(0.2, _world())
    assert pl.last_macro_action == 2
    assert pl.last_return_values[2] == pytest.approx(50.0, abs=1.0)
    assert pl.last_return_values[0] == pytest.approx(-500.0, abs=1.0)
    # one virtual visit (the default) is washed out by the search: values move off the prior
    pl = _cpu(BehaviorUCTNHeuristicRiskConstraint, _params(256, wave=8), _Prior(values=values))
    pl.Plan(0.2, _world())
    # (every action the search actually visits: here all but the first, which the prior rules out)
    moved = [abs(pl.last_return_values[a] - values["Return"][a]) for a in range(4)]
    assert sorted(moved)[1] > 20.0 and max(moved) > 400.0


def test_root_parallel_trees_share_the_model():
    """Mcts::NumParallelMcts trees on threads: calls into the model are serialised"""
    prior = _Prior(policy=[1, 1, 1, 1])
    p = _params(32, wave=8, **{W.P0 + "Mcts::NumParallelMcts": 4, W.P0 + "Mcts::UseMultiThreading": 1})
    pl = _cpu(BehaviorUCTNHeuristicRiskConstraint, p, prior)
    pl.Plan(0.2, _world())
    assert pl.last_num_iterations == 4 * 32
    assert pl.node_prior_stats[0] == len(prior.batches) == 4 * (1 + 32 // 8)


def test_empty_prior_is_the_plain_risk_constraint_planner():
    a = _cpu(BehaviorUCTNHeuristicRiskConstraint, _params(128, wave=8), _Prior())
    b = _cpu(P.BehaviorUCTRiskConstraint, _params(128, wave=8))
    world = _world()
    a.Plan(0.2, world)
    b.Plan(0.2, world)
    assert np.array_equal(a.root_stats(), b.root_stats())
    assert a.last_macro_action == b.last_macro_action


def test_model_errors_surface_and_a_missing_prior_is_refused():
    def broken(obs):
        raise RuntimeError("model exploded")
    pl = _cpu(BehaviorUCTNHeuristicRiskConstraint, _params(8), broken)
    with pytest.raises(RuntimeError, match="model exploded"):
        pl.Plan(0.2, _world())
    # native layer: the kind without a callback cannot plan
    lib = P.declare_planner_api(oracle.planner_lib())
    h = C.c_void_p()
    assert lib.oracle_planner_create(P.PLANNER_NHEURISTIC_RISK_CONSTRAINT, C.byref(h)) == _abi.OK
    world = _world()
    assert lib.bmcts_planner_set_map(h, C.byref(world.map)) == _abi.OK
    arr = (P.AgentState * 1)(P.AgentState(1, 5.0, -1.75, 0.0, 5.0, 4.0, 2.0, 1.25, 0.0))
    res = P.PlanResult()
    assert lib.bmcts_planner_plan(h, 0.2, 0.0, 1, arr, 1, C.byref(res)) != _abi.OK
    assert b"set_node_prior" in lib.bmcts_planner_last_error(h)
    lib.bmcts_planner_destroy(h)


def _torch_model(n_obs, n_actions, best):
    import torch

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.body = torch.nn.Linear(n_obs, 3 * n_actions)
            with torch.no_grad():
                self.body.weight.zero_()
                self.body.bias.zero_()
                self.body.bias[2 * n_actions:] = -300.0          # returns
                self.body.bias[2 * n_actions + best] = 40.0
                self.body.weight[0, 3] = 1.0                      # envelope risk of action 0 = ego v

        def forward(self, x):
            return self.body(x)
    return Net()


def test_torch_model_through_the_converter(tmp_path):
    import torch
    conv = NNToValueConverterSequential(4)
    model = _torch_model(20, 4, best=1)
    prior = TorchNodePrior(model, conv, "cpu")
    out = prior(np.full((3, 20), 0.25, np.float32))
    assert out["Return"].shape == (3, 4) and out["Return"][0, 1] == 40.0
    assert out["EnvelopeRisk"][2, 0] == pytest.approx(0.25)
    # TorchScript file, as the reference's constructor takes it
    path = str(tmp_path / "model.pt")
    torch.jit.script(model).save(path)
    heavy = _params(16, wave=4, **{W.P0 + "Mcts::NeuralStatistic::PriorVisits": 100000})
    pl = _cpu(BehaviorUCTNHeuristicRiskConstraint, heavy, path, None, conv)
    pl.Plan(0.2, _world())
    assert pl.GetModelFileName() == path and pl.GetNNToValueConverter() is conv
    assert pl.last_macro_action == 1
    assert pl.node_prior_stats[0] == 1 + 16 // 4


@pytest.mark.gpu
@pytest.mark.skipif(not has_gpu(), reason="needs a CUDA device")
def test_gpu_planner_with_a_model_on_the_gpu():
    """product path: rollouts on the CUDA engine, the model evaluated on the same device, one
    batch per wave"""
    conv = NNToValueConverterSequential(4)
    heavy = _params(256, wave=32, **{W.P0 + "Mcts::NeuralStatistic::PriorVisits": 100000})
    pl = BehaviorUCTNHeuristicRiskConstraint(heavy, W.hypotheses(500), None,
                                             _torch_model(20, 4, best=3), None, conv, device=0)
    pl.Plan(0.2, _world())
    assert pl.last_macro_action == 3 and pl.last_num_iterations == 256
    assert pl.node_prior_stats[0] == 1 + 256 // 32
    assert pl._prior.device.type == "cuda" and pl._prior.batches == 9
    assert pl.engine_launches > 0
