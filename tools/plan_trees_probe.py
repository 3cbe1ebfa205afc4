#!/usr/bin/env python
"""Plan() iterations/s against the number of root-parallel trees per host thread (interleaved
waves), at the bench's rollout horizon and at the reference's default cap of 1000 steps -- run
on the GPU box."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2_merge_sbg"
threads = min(16, os.cpu_count() or 1)
for horizon, iterations in ((20, 32768), (1000, 4096)):
    for per_thread in (1, 2, 4, 8):
        r = bench.plan_benchmark(wl, horizon, iterations, 0, "gpu", trees=per_thread * threads,
                                 threads=threads)
        print(f"{wl} horizon {horizon}: {per_thread} trees/thread x {threads} threads: "
              f"{r['iterations_per_s']:.4g} iterations/s, Plan() {r['plan_ms']:.1f} ms", flush=True)
