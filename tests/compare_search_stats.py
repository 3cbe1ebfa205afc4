#!/usr/bin/env python
"""Refactoring guard for the host tree search: the root statistics of a set of searches
(three planners x sequential / wave-batched / multi-tree) from two builds of the oracle-backed
planner library must be bit-identical.

    git archive <old commit> | tar -x -C /tmp/old && make -C /tmp/old/oracle
    python tests/compare_search_stats.py /tmp/old/oracle/_build/liboracle_planner.so \
        oracle/_build/liboracle_planner.so
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))  # test infrastructure: uses the CPU oracle
import oracle, planner_worlds as W
from planner_mcts_b200 import planner as P
def run(libpath):
    lib = C.CDLL(libpath)
    out = []
    for cls, extra in ((P.BehaviorUCTHypothesis, {}), (P.BehaviorUCTRiskConstraint, {W.P0+"Mcts::State::SplitSafeDistCollision":1, W.P0+"Mcts::State::EvaluationParameters::AddSafeDist":1, "BehaviorUctRiskConstraint::DefaultAvailableRisk":[0.1,0.0]}), (P.BehaviorUCTCooperative, {})):
        for wave, trees in ((1,1),(16,1),(8,5)):
            params = W.base_params(600); params.update(extra)
            params[W.P0+"Mcts::Engine::WaveSize"] = wave
            params[W.P0+"Mcts::NumParallelMcts"] = trees
            world = W.observed_world(2, 4.0, 6.0, 3.0, goal_center=(150.0, -1.75))
            kw = dict(_lib=lib, _create=lambda kind, o: lib.oracle_planner_create(kind, o))
            pl = cls(params, **kw) if cls is P.BehaviorUCTCooperative else cls(params, W.hypotheses(300), **kw)
            pl.Plan(0.2, world); pl.Plan(0.2, world)
            out.append((cls.__name__, wave, trees, pl.root_stats(), pl.last_macro_action, pl.last_num_nodes))
    return out
a = run(sys.argv[1])
b = run(sys.argv[2])
ok = True
for x, y in zip(a, b):
    same = np.array_equal(x[3], y[3]) and x[4:] == y[4:]
    ok &= same
    print(x[0], x[1], x[2], "identical" if same else "DIFFERENT", x[4:], y[4:])
print("ALL IDENTICAL" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
