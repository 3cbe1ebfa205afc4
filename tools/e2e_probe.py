#!/usr/bin/env python
"""Where the end-to-end time of a host-memory rollout batch goes (run on the GPU box)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from planner_mcts_b200 import RolloutEngine, _abi, scenarios  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
cfg, scn, meta = scenarios.workload("c2_merge_sbg", horizon=20)
pose, alive, hyp = scenarios.leaves_for(scn, meta, n, seed=1)
eng = RolloutEngine(cfg, scn)
dev = torch.device("cuda", 0)
h_pose = torch.from_numpy(pose).pin_memory()
h_alive = torch.from_numpy(alive.view(np.int32)).pin_memory()
h_hyp = torch.from_numpy(hyp).pin_memory()
h_res = torch.zeros(n, 8, dtype=torch.float32).pin_memory()
d_pose = torch.empty_like(h_pose, device=dev)
d_res = torch.zeros(n, 8, dtype=torch.float32, device=dev)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


print("H2D pose %.1f MB: %.2f ms" % (h_pose.numel() * 4 / 1e6, timed(lambda: d_pose.copy_(h_pose, non_blocking=True))))
print("D2H result %.1f MB: %.2f ms" % (h_res.numel() * 4 / 1e6, timed(lambda: h_res.copy_(d_res, non_blocking=True))))
d_alive = torch.from_numpy(alive.view(np.int32)).to(dev)
d_hyp = torch.from_numpy(hyp).to(dev)
print("kernel device-resident: %.2f ms" % timed(lambda: (eng.rollout_raw(n, _abi.MEM_DEVICE, d_pose, d_alive, d_hyp, None, d_res), eng.sync())))
for chunk in (16384, 32768, 65536, 131072):
    os.environ["BMCTS_STREAM_CHUNK"] = str(chunk)
    ms = timed(lambda: eng.rollout_raw(n, _abi.MEM_HOST, h_pose, h_alive, h_hyp, None, h_res))
    print("host call, streamed chunk %7d: %.2f ms" % (chunk, ms))
os.environ["BMCTS_STREAM_CHUNK"] = "0"
for chunk in (0, 262144):
    os.environ["BMCTS_PIPE_CHUNK"] = str(chunk)
    ms = timed(lambda: eng.rollout_raw(n, _abi.MEM_HOST, h_pose, h_alive, h_hyp, None, h_res))
    print("host call, chunk %7d: %.2f ms" % (chunk, ms))
t0 = time.perf_counter()
s = int(h_res.numpy().view(np.int32)[:, 6].sum())
print("host sum of agent_steps: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
