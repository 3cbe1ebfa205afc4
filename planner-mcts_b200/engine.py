"""Python host binding of the rollout engine (ctypes over the C ABI).

`RolloutEngine` is the leaf evaluator the planners plug in where the reference
instantiates `mcts::RandomHeuristic` (behavior_uct_hypothesis.cpp:59-60,
behavior_uct_risk_constraint.cpp:47-48, behavior_uct_cooperative.cpp:41-42) and
where they call `MctsState*::execute` for expansions.  numpy arrays travel as
host buffers (copies inside the call); torch CUDA tensors travel as device
pointers (no copies, stream-ordered on the engine's stream).

All batch arrays are agent-major: pose[A, n, 4], hyp_id[A, n], alive[n] bitmask.
"""
import ctypes as C

import numpy as np

from . import _abi
from ._abi import BmctsError

RESULT_DTYPE = np.dtype([
    ("ret_ego", np.float32), ("cost0", np.float32), ("cost1", np.float32),
    ("step_len", np.float32), ("flags", np.uint32), ("n_steps", np.int32),
    ("agent_steps", np.int32), ("final_alive", np.uint32)])
assert RESULT_DTYPE.itemsize == C.sizeof(_abi.RolloutResult)


def _host(arr, dtype, shape=None, name="array"):
    if arr is None:
        return None
    a = np.ascontiguousarray(arr, dtype=dtype)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(a.shape)}")
    return a


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()  # torch tensor


class RolloutEngine:
    """One engine = one CUDA stream on one GPU (bmcts_create)."""

    def __init__(self, config, scenario, device=0):
        self._lib = _abi.load_library()
        self._h = C.c_void_p()
        self.config = config
        self.scenario = scenario
        st = self._lib.bmcts_create(int(device), C.byref(config), C.byref(self._h))
        if st != _abi.OK:
            self._h = C.c_void_p()
            raise BmctsError(st, self._lib.bmcts_status_string(st).decode())
        self._check(self._lib.bmcts_set_scenario(self._h, C.byref(scenario)))
        self.A = scenario.n_agents
        self.H = scenario.n_hypotheses
        self.device = int(device)

    # -- plumbing ----------------------------------------------------------
    def _check(self, st):
        if st != _abi.OK:
            msg = self._lib.bmcts_last_error(self._h).decode()
            raise BmctsError(st, msg or self._lib.bmcts_status_string(st).decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.bmcts_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def math_probe(self, op, a, b):
        """out[i] = f(a[i], b[i]) of include/bmcts_math.h evaluated on the device (self test)"""
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        out = np.zeros_like(a)
        fp = C.POINTER(C.c_float)
        self._lib.bmcts_math_probe.argtypes = [C.c_void_p, C.c_int, fp, fp, C.c_int, fp]
        self._check(self._lib.bmcts_math_probe(self._h, _abi.MATH_OPS.index(op), a.ctypes.data_as(fp),
                                               b.ctypes.data_as(fp), a.size, out.ctypes.data_as(fp)))
        return out

    def set_config(self, config):
        self._check(self._lib.bmcts_set_config(self._h, C.byref(config)))
        self.config = config

    def sync(self):
        self._check(self._lib.bmcts_sync(self._h))

    def last_kernel_ms(self):
        ms = C.c_float()
        self._check(self._lib.bmcts_last_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def set_kernel_timing(self, on):
        """event records around every launch on / off (on by default; off: last_kernel_ms raises)"""
        self._check(self._lib.bmcts_set_kernel_timing(self._h, 1 if on else 0))

    def last_launch_geometry(self):
        """(lanes per rollout, rollout warps per block, ahead variant?) of the last kernel launch"""
        t, w, a = C.c_int(), C.c_int(), C.c_int()
        self._check(self._lib.bmcts_last_launch_geometry(self._h, C.byref(t), C.byref(w), C.byref(a)))
        return int(t.value), int(w.value), bool(a.value)

    def launch_count(self):
        return int(self._lib.bmcts_launch_count(self._h))

    def stream_ptr(self):
        return self._lib.bmcts_stream(self._h)

    # -- rollout (RandomHeuristic::calculate_heuristic_values) --------------
    def rollout(self, pose, alive, hyp_id=None, depth=None, seed=1000, stream=0,
                first_rollout_id=0, want_agents=False, want_final=False, state_index=None):
        """Host-buffer rollout batch.  Returns a dict of numpy arrays.  With state_index[n],
        pose / alive / depth are tables of distinct states and rollout k starts from row
        state_index[k] (hyp_id stays per rollout); pose is given agent-major [A, n_states, 4] like
        every batch array and handed to the C ABI state-major, as its unique-state form asks."""
        A = self.A
        pose = _host(pose, np.float32, name="pose")
        ns = pose.shape[1]
        n = ns if state_index is None else len(state_index)
        pose = _host(pose, np.float32, (A, ns, 4), "pose")
        if state_index is not None:
            pose = np.ascontiguousarray(pose.transpose(1, 0, 2))   # [n_states, A, 4]
        alive = _host(alive, np.uint32, (ns,), "alive")
        hyp_id = _host(hyp_id if hyp_id is not None else np.zeros((A, n)), np.uint8, (A, n), "hyp_id")
        depth = _host(depth, np.uint16, (ns,), "depth")
        state_index = _host(state_index, np.uint32, (n,), "state_index")
        res = np.zeros(n, dtype=RESULT_DTYPE)
        ret_agents = np.zeros((A, n), np.float32) if want_agents else None
        final_pose = np.zeros((A, n, 4), np.float32) if want_final else None
        inb = _abi.LeafBatch(n, _abi.MEM_HOST, _ptr(pose), _ptr(alive), _ptr(hyp_id), _ptr(depth),
                             int(seed), int(stream), 0, int(first_rollout_id),
                             0 if state_index is None else ns, 0, _ptr(state_index))
        out = _abi.RolloutOut(_ptr(res), _ptr(ret_agents), _ptr(final_pose))
        self._check(self._lib.bmcts_rollout_batch(self._h, C.byref(inb), C.byref(out)))
        d = {"result": res}
        if want_agents:
            d["ret_agents"] = ret_agents
        if want_final:
            d["final_pose"] = final_pose
        return d

    def rollout_raw(self, n, memory, pose, alive, hyp_id, depth, result, ret_agents=None,
                    final_pose=None, seed=1000, stream=0, first_rollout_id=0, n_states=0,
                    state_index=None):
        """Rollout batch over caller-owned buffers (numpy arrays, torch tensors or
        integer addresses); memory = MEM_HOST or MEM_DEVICE.  Asynchronous for
        MEM_DEVICE (call sync())."""
        def p(x):
            return x if (x is None or isinstance(x, int)) else _ptr(x)
        inb = _abi.LeafBatch(int(n), int(memory), p(pose), p(alive), p(hyp_id), p(depth), int(seed),
                             int(stream), 0, int(first_rollout_id), int(n_states), 0,
                             p(state_index))
        out = _abi.RolloutOut(p(result), p(ret_agents), p(final_pose))
        self._check(self._lib.bmcts_rollout_batch(self._h, C.byref(inb), C.byref(out)))

    # -- step (MctsState*::execute + plan_action_current_hypothesis) --------
    def _step_buffers(self, pose, alive, hyp_id, depth, action, draw_id):
        A = self.A
        pose = _host(pose, np.float32, name="pose")
        n = pose.shape[1]
        pose = _host(pose, np.float32, (A, n, 4), "pose")
        alive = _host(alive, np.uint32, (n,), "alive")
        hyp_id = _host(hyp_id if hyp_id is not None else np.zeros((A, n)), np.uint8, (A, n), "hyp_id")
        depth = _host(depth, np.uint16, (n,), "depth")
        action = _host(action, np.uint8, (A, n), "action")
        draw_id = _host(draw_id if draw_id is not None else np.zeros((A, n)), np.uint64, (A, n),
                        "draw_id")
        return n, pose, alive, hyp_id, depth, action, draw_id

    def step(self, pose, alive, hyp_id, action, draw_id=None, depth=None, seed=1000, stream=0):
        A = self.A
        n, pose, alive, hyp_id, depth, action, draw_id = self._step_buffers(
            pose, alive, hyp_id, depth, action, draw_id)
        o = {
            "next_pose": np.zeros((A, n, 4), np.float32), "next_alive": np.zeros(n, np.uint32),
            "reward": np.zeros((A, n), np.float32), "cost": np.zeros((2, n), np.float32),
            "flags": np.zeros(n, np.uint32), "last_action": np.zeros((A, n), np.float32)}
        inb = _abi.StepIn(n, _abi.MEM_HOST, _ptr(pose), _ptr(alive), _ptr(hyp_id), _ptr(depth),
                          _ptr(action), _ptr(draw_id), int(seed), int(stream), 0)
        out = _abi.StepOut(_ptr(o["next_pose"]), _ptr(o["next_alive"]), _ptr(o["reward"]),
                           _ptr(o["cost"]), _ptr(o["flags"]), _ptr(o["last_action"]))
        self._check(self._lib.bmcts_step_batch(self._h, C.byref(inb), C.byref(out)))
        return o

    def expand_rollout(self, pose, alive, hyp_id, action, draw_id=None, depth=None, seed=1000,
                       stream=0, first_rollout_id=0, want_agents=False, want_final=False,
                       _async=False):
        A = self.A
        n, pose, alive, hyp_id, depth, action, draw_id = self._step_buffers(
            pose, alive, hyp_id, depth, action, draw_id)
        o = {
            "next_pose": np.zeros((A, n, 4), np.float32), "next_alive": np.zeros(n, np.uint32),
            "reward": np.zeros((A, n), np.float32), "cost": np.zeros((2, n), np.float32),
            "flags": np.zeros(n, np.uint32), "last_action": np.zeros((A, n), np.float32),
            "result": np.zeros(n, dtype=RESULT_DTYPE)}
        if want_agents:
            o["ret_agents"] = np.zeros((A, n), np.float32)
        if want_final:
            o["final_pose"] = np.zeros((A, n, 4), np.float32)
        inb = _abi.StepIn(n, _abi.MEM_HOST, _ptr(pose), _ptr(alive), _ptr(hyp_id), _ptr(depth),
                          _ptr(action), _ptr(draw_id), int(seed), int(stream), 0)
        so = _abi.StepOut(_ptr(o["next_pose"]), _ptr(o["next_alive"]), _ptr(o["reward"]),
                          _ptr(o["cost"]), _ptr(o["flags"]), _ptr(o["last_action"]))
        ro = _abi.RolloutOut(_ptr(o["result"]), _ptr(o.get("ret_agents")), _ptr(o.get("final_pose")))
        if _async:
            # the two halves of the call: the arrays must stay alive until expand_rollout_end()
            self._check(self._lib.bmcts_expand_rollout_batch_begin(
                self._h, C.byref(inb), C.byref(so), int(first_rollout_id), C.byref(ro)))
            self._pending = (o, pose, alive, hyp_id, depth, action, draw_id)
            return None
        self._check(self._lib.bmcts_expand_rollout_batch(self._h, C.byref(inb), C.byref(so),
                                                         int(first_rollout_id), C.byref(ro)))
        return o

    def expand_rollout_begin(self, *args, **kw):
        """bmcts_expand_rollout_batch_begin: enqueue and return; collect with expand_rollout_end()"""
        return self.expand_rollout(*args, _async=True, **kw)

    def expand_rollout_end(self):
        self._check(self._lib.bmcts_expand_rollout_batch_end(self._h))
        o, self._pending = self._pending[0], None
        return o

    # -- likelihood (BehaviorHypothesisIDM::GetProbability) -----------------
    def likelihood(self, pose, alive, action, n_samples=10000, n_buckets=100, lower=-10.0,
                   upper=10.0, seed=1000, stream=0):
        """prob[A, H, n_worlds]"""
        A, H = self.A, self.H
        pose = _host(pose, np.float32, name="pose")
        W = pose.shape[1]
        pose = _host(pose, np.float32, (A, W, 4), "pose")
        alive = _host(alive, np.uint32, (W,), "alive")
        action = _host(action, np.float32, (A, W), "action")
        prob = np.zeros((A, H, W), np.float32)
        inb = _abi.LikelihoodIn(W, _abi.MEM_HOST, _ptr(pose), _ptr(alive), _ptr(action),
                                int(n_samples), int(n_buckets), float(lower), float(upper),
                                int(seed), int(stream), 0)
        self._check(self._lib.bmcts_likelihood_batch(self._h, C.byref(inb), _ptr(prob)))
        return prob

    def likelihood_raw(self, n_worlds, memory, pose, alive, action, prob, n_samples=10000,
                       n_buckets=100, lower=-10.0, upper=10.0, seed=1000, stream=0):
        """Likelihood batch over caller-owned buffers (numpy arrays, torch tensors or integer
        addresses); asynchronous for MEM_DEVICE (call sync())."""
        def p(x):
            return x if (x is None or isinstance(x, int)) else _ptr(x)
        inb = _abi.LikelihoodIn(int(n_worlds), int(memory), p(pose), p(alive), p(action),
                                int(n_samples), int(n_buckets), float(lower), float(upper),
                                int(seed), int(stream), 0)
        self._check(self._lib.bmcts_likelihood_batch(self._h, C.byref(inb), p(prob)))
