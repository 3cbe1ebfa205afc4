#!/usr/bin/env python
"""One MCTS-sized wave (15 leaves, expansion + 20-step rollouts) per call, a few calls: the
launch the ego-ahead variants are made for.  Run plainly for the kernel / call latency, or under
ncu (`-k regex:rollout_lane -s 3 -c 1`) for a capture of `rollout_lane_kernel<16, true>`.
usage: python tools/small_wave_probe.py [workload] [n_leaves]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from planner_mcts_b200 import RolloutEngine, scenarios  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2_merge_sbg"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 15
cfg, scn, meta = scenarios.workload(wl, horizon=20)
eng = RolloutEngine(cfg, scn)
pose, alive, hyp = scenarios.leaves_for(scn, meta, n, seed=3)
action = np.zeros((pose.shape[0], n), dtype=np.uint8)
ks, ws = [], []
for r in range(12):
    t0 = time.perf_counter()
    eng.expand_rollout(pose, alive, hyp, action, seed=5, stream=r)
    ws.append((time.perf_counter() - t0) * 1e6)
    ks.append(eng.last_kernel_ms() * 1e3)
t, w, ahead = eng.last_launch_geometry()
print(f"{wl} n={n}: {t} lanes per rollout, ahead={ahead}: kernel {np.median(ks[3:]):.1f} us, "
      f"call {np.median(ws[3:]):.1f} us")
