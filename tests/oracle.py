"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so) for the tests.

The oracle is the checker: tests compare the CUDA path against it.  It takes the
same PODs as the C ABI (planner_mcts_b200._abi mirrors include/bmcts_c.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from planner_mcts_b200 import _abi
from planner_mcts_b200.engine import RESULT_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "_build", "liboracle.so")

_lib = None


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
    return ORACLE_LIB


def load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = C.CDLL(ORACLE_LIB)
    P = C.POINTER
    u32p = P(C.c_uint32)
    lib.oracle_philox4x32_10.argtypes = [u32p, u32p, u32p]
    lib.oracle_philox4x32_10.restype = None
    lib.oracle_prediction_time_span.argtypes = [P(_abi.Config), C.c_uint]
    lib.oracle_prediction_time_span.restype = C.c_float
    lib.oracle_reward_from_flags.argtypes = [P(_abi.Config), C.c_uint32, C.c_float]
    lib.oracle_reward_from_flags.restype = C.c_float
    lib.oracle_costs_from_flags.argtypes = [P(_abi.Config), C.c_uint32, C.c_float, P(C.c_float)]
    lib.oracle_terminal_from_flags.argtypes = [P(_abi.Config), C.c_uint32]
    lib.oracle_terminal_from_flags.restype = C.c_uint32
    lib.oracle_project.argtypes = [P(_abi.Corridor), C.c_float, C.c_float, P(C.c_float),
                                   P(C.c_float), P(C.c_int), P(C.c_int)]
    lib.oracle_project.restype = None
    lib.oracle_rollout_batch.argtypes = [P(_abi.Config), P(_abi.Scenario), P(_abi.LeafBatch),
                                         P(_abi.RolloutOut), C.c_int]
    lib.oracle_step_batch.argtypes = [P(_abi.Config), P(_abi.Scenario), P(_abi.StepIn),
                                      P(_abi.StepOut)]
    lib.oracle_expand_rollout_batch.argtypes = [P(_abi.Config), P(_abi.Scenario), P(_abi.StepIn),
                                                P(_abi.StepOut), C.c_uint64, P(_abi.RolloutOut),
                                                C.c_int]
    lib.oracle_likelihood_batch.argtypes = [P(_abi.Config), P(_abi.Scenario),
                                            P(_abi.LikelihoodIn), C.c_void_p]
    _lib = lib
    return lib


def math_probe(op, a, b):
    lib = load()
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    out = np.zeros_like(a)
    fp = C.POINTER(C.c_float)
    lib.oracle_math_probe.argtypes = [C.c_int, fp, fp, C.c_int, fp]
    lib.oracle_math_probe.restype = None
    lib.oracle_math_probe(_abi.MATH_OPS.index(op), a.ctypes.data_as(fp), b.ctypes.data_as(fp), a.size,
                          out.ctypes.data_as(fp))
    return out


def _h(a, dtype, shape, name):
    a = np.ascontiguousarray(a, dtype=dtype)
    assert tuple(a.shape) == tuple(shape), f"{name}: {a.shape} != {shape}"
    return a


def _p(a):
    return None if a is None else a.ctypes.data


def _check(st):
    if st != 0:
        raise RuntimeError(f"oracle status {st}")


def philox(ctr, key):
    lib = load()
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib.oracle_philox4x32_10(c, k, o)
    return list(o)


def rollout(cfg, scn, pose, alive, hyp_id=None, depth=None, seed=1000, stream=0,
            first_rollout_id=0, n_threads=1, state_index=None):
    lib = load()
    A = scn.n_agents
    ns = np.asarray(pose).shape[1]
    n = ns if state_index is None else len(state_index)
    pose = _h(pose, np.float32, (A, ns, 4), "pose")
    if state_index is not None:   # the unique-state form takes the table state-major
        pose = np.ascontiguousarray(pose.transpose(1, 0, 2))
    alive = _h(alive, np.uint32, (ns,), "alive")
    hyp_id = _h(hyp_id if hyp_id is not None else np.zeros((A, n)), np.uint8, (A, n), "hyp_id")
    depth = None if depth is None else _h(depth, np.uint16, (ns,), "depth")
    state_index = None if state_index is None else _h(state_index, np.uint32, (n,), "state_index")
    res = np.zeros(n, dtype=RESULT_DTYPE)
    ret_agents = np.zeros((A, n), np.float32)
    final_pose = np.zeros((A, n, 4), np.float32)
    inb = _abi.LeafBatch(n, _abi.MEM_HOST, _p(pose), _p(alive), _p(hyp_id), _p(depth), int(seed),
                         int(stream), 0, int(first_rollout_id),
                         0 if state_index is None else ns, 0, _p(state_index))
    out = _abi.RolloutOut(_p(res), _p(ret_agents), _p(final_pose))
    _check(lib.oracle_rollout_batch(C.byref(cfg), C.byref(scn), C.byref(inb), C.byref(out),
                                    int(n_threads)))
    return {"result": res, "ret_agents": ret_agents, "final_pose": final_pose}


def _step_in(scn, pose, alive, hyp_id, depth, action, draw_id, seed, stream):
    A = scn.n_agents
    n = np.asarray(pose).shape[1]
    pose = _h(pose, np.float32, (A, n, 4), "pose")
    alive = _h(alive, np.uint32, (n,), "alive")
    hyp_id = _h(hyp_id if hyp_id is not None else np.zeros((A, n)), np.uint8, (A, n), "hyp_id")
    depth = None if depth is None else _h(depth, np.uint16, (n,), "depth")
    action = _h(action, np.uint8, (A, n), "action")
    draw_id = _h(draw_id if draw_id is not None else np.zeros((A, n)), np.uint64, (A, n), "draw_id")
    keep = (pose, alive, hyp_id, depth, action, draw_id)
    inb = _abi.StepIn(n, _abi.MEM_HOST, _p(pose), _p(alive), _p(hyp_id), _p(depth), _p(action),
                      _p(draw_id), int(seed), int(stream), 0)
    return n, inb, keep


def _step_out(A, n):
    o = {"next_pose": np.zeros((A, n, 4), np.float32), "next_alive": np.zeros(n, np.uint32),
         "reward": np.zeros((A, n), np.float32), "cost": np.zeros((2, n), np.float32),
         "flags": np.zeros(n, np.uint32), "last_action": np.zeros((A, n), np.float32)}
    so = _abi.StepOut(_p(o["next_pose"]), _p(o["next_alive"]), _p(o["reward"]), _p(o["cost"]),
                      _p(o["flags"]), _p(o["last_action"]))
    return o, so


def step(cfg, scn, pose, alive, hyp_id, action, draw_id=None, depth=None, seed=1000, stream=0):
    lib = load()
    n, inb, keep = _step_in(scn, pose, alive, hyp_id, depth, action, draw_id, seed, stream)
    o, so = _step_out(scn.n_agents, n)
    _check(lib.oracle_step_batch(C.byref(cfg), C.byref(scn), C.byref(inb), C.byref(so)))
    return o


def expand_rollout(cfg, scn, pose, alive, hyp_id, action, draw_id=None, depth=None, seed=1000,
                   stream=0, first_rollout_id=0):
    lib = load()
    A = scn.n_agents
    n, inb, keep = _step_in(scn, pose, alive, hyp_id, depth, action, draw_id, seed, stream)
    o, so = _step_out(A, n)
    o["result"] = np.zeros(n, dtype=RESULT_DTYPE)
    o["ret_agents"] = np.zeros((A, n), np.float32)
    o["final_pose"] = np.zeros((A, n, 4), np.float32)
    ro = _abi.RolloutOut(_p(o["result"]), _p(o["ret_agents"]), _p(o["final_pose"]))
    _check(lib.oracle_expand_rollout_batch(C.byref(cfg), C.byref(scn), C.byref(inb), C.byref(so),
                                           int(first_rollout_id), C.byref(ro), 1))
    return o


def likelihood(cfg, scn, pose, alive, action, n_samples=10000, n_buckets=100, lower=-10.0,
               upper=10.0, seed=1000, stream=0):
    lib = load()
    A, H = scn.n_agents, scn.n_hypotheses
    W = np.asarray(pose).shape[1]
    pose = _h(pose, np.float32, (A, W, 4), "pose")
    alive = _h(alive, np.uint32, (W,), "alive")
    action = _h(action, np.float32, (A, W), "action")
    prob = np.zeros((A, H, W), np.float32)
    inb = _abi.LikelihoodIn(W, _abi.MEM_HOST, _p(pose), _p(alive), _p(action), int(n_samples),
                            int(n_buckets), float(lower), float(upper), int(seed), int(stream), 0)
    _check(lib.oracle_likelihood_batch(C.byref(cfg), C.byref(scn), C.byref(inb), _p(prob)))
    return prob


# ---- CPU-backed planner (the host tree search of planner-mcts_b200/host on top of the oracle)

ORACLE_PLANNER_LIB = os.path.join(ORACLE_DIR, "_build", "liboracle_planner.so")
_planner_lib = None


def planner_lib():
    """liboracle_planner.so: same planner C ABI as libbmcts.so, leaf evaluation by the oracle"""
    global _planner_lib
    if _planner_lib is None:
        build()
        _planner_lib = C.CDLL(ORACLE_PLANNER_LIB)
        _planner_lib.oracle_planner_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    return _planner_lib


def make_planner(cls, params=None, hypotheses=None):
    """cls = planner_mcts_b200.planner.BehaviorUCT*; returns the CPU-backed twin"""
    lib = planner_lib()
    kw = dict(params=params, _lib=lib, _create=lambda kind, out: lib.oracle_planner_create(kind, out))
    if hypotheses is not None:
        kw["hypotheses"] = hypotheses
    return cls(**kw)
