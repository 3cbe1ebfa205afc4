#!/usr/bin/env python
"""K4 (BehaviorHypothesisIDM::GetProbability as one batched kernel): GPU vs CPU oracle time for
the belief update of one Plan() (all agents x all hypotheses x NumSamples) -- run on the GPU box."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from planner_mcts_b200 import RolloutEngine, scenarios  # noqa: E402

for name, worlds in (("c2_merge_sbg", 1), ("c5_rsbg", 1), ("c5_rsbg", 16)):
    cfg, scn, meta = scenarios.workload(name)
    pose, alive, hyp = scenarios.leaves_for(scn, meta, worlds, seed=11)
    rng = np.random.default_rng(0)
    action = rng.uniform(-3, 2, size=(scn.n_agents, worlds)).astype(np.float32)
    eng = RolloutEngine(cfg, scn)
    g = eng.likelihood(pose, alive, action, n_samples=10000, n_buckets=100)
    t0 = time.perf_counter()
    for _ in range(10):
        g = eng.likelihood(pose, alive, action, n_samples=10000, n_buckets=100)
    gpu_ms = (time.perf_counter() - t0) / 10 * 1e3
    t0 = time.perf_counter()
    c = oracle.likelihood(cfg, scn, pose, alive, action, n_samples=10000, n_buckets=100)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    evals = worlds * scn.n_agents * scn.n_hypotheses * 10000
    print(f"{name} worlds={worlds}: {evals:.2e} IDM evaluations; GPU {gpu_ms:.3f} ms (host call, "
          f"kernel {eng.last_kernel_ms():.3f} ms), CPU oracle 1 thread {cpu_ms:.1f} ms, "
          f"equal={np.array_equal(g, c)}")
