#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the CPU oracle.

These are REGRESSION fixtures of the normative step spec, not outputs of
the reference: bark-simulator/planner-mcts cannot be built or imported in this environment (it
needs Bazel + BARK + mamcts), and its own tests hold no numeric
trajectories.  They pin today's behaviour of the oracle AND of the shared elementary math
(include/bmcts_math.h), so that a change to either -- which the oracle-vs-CUDA parity tests alone
would not notice, both sides moving together -- shows up as a diff of committed numbers.
Regenerate only for a deliberate spec change and say so in the commit:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle  # noqa: E402
from planner_mcts_b200 import scenarios  # noqa: E402

WORKLOADS = ["c1_highway_mdp", "c2_merge_sbg", "c3_merge_risk", "c4_coop", "c5_rsbg",
             "x_curved_max"]
SEED, STREAM, FIRST_ID = 1000, 2, 4242
# Euler step of the ego's macro action: the reference default (0.02 s = 25 steps per 0.5 s,
# tests/data/nheuristic_test.json:153-155) on three fixtures, the benchmark setting of
# SURVEY.md 8d (0.05 s = 10 steps) on the others
EGO_DT = {"c1_highway_mdp": 0.02, "c3_merge_risk": 0.02, "c4_coop": 0.02}


def inputs(name):
    cfg, scn, meta = scenarios.workload(name, horizon=20, ego_dt=EGO_DT.get(name, 0.05))
    rng = np.random.default_rng(99)
    pose, alive, hyp = scenarios.sample_leaves(scn, 24, meta["lane_of"], meta["s_of"], rng,
                                               dead_prob=0.08)
    A = scn.n_agents
    action = rng.integers(0, scn.n_ego_actions, size=(A, 24)).astype(np.uint8)
    draw = rng.integers(0, 2**62, size=(A, 24)).astype(np.uint64)
    depth = rng.integers(1, 5, size=24).astype(np.uint16)
    return cfg, scn, dict(pose=pose, alive=alive, hyp=hyp, action=action, draw=draw, depth=depth)


def compute(name, rollout_fn, expand_fn):
    cfg, scn, x = inputs(name)
    r = rollout_fn(cfg, scn, x["pose"], x["alive"], x["hyp"], depth=x["depth"], seed=SEED,
                   stream=STREAM, first_rollout_id=FIRST_ID)
    e = expand_fn(cfg, scn, x["pose"], x["alive"], x["hyp"], x["action"], x["draw"],
                  depth=x["depth"], seed=SEED, stream=STREAM, first_rollout_id=FIRST_ID)
    out = {"rollout_result": r["result"], "rollout_ret_agents": r["ret_agents"],
           "rollout_final_pose": r["final_pose"]}
    for k in ("next_pose", "next_alive", "reward", "cost", "flags", "last_action"):
        out["step_" + k] = e[k]
    out["expand_result"] = e["result"]
    return out


def path(name):
    return os.path.join(HERE, f"{name}.npz")


if __name__ == "__main__":
    for name in (sys.argv[1:] or WORKLOADS):   # optionally only the named fixtures
        out = compute(name, oracle.rollout, oracle.expand_rollout)
        np.savez_compressed(path(name), **out)
        print(name, {k: v.shape for k, v in out.items()})
