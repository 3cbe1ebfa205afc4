#!/usr/bin/env python
"""Plan() time split (host select / fill / engine / backup per tree) at several wave sizes and
tree counts -- run on the GPU box."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["BMCTS_DEBUG"] = "1"
import bench  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2_merge_sbg"
for trees, wave in ((1, 1024), (1, 4096), (8, 1024), (8, 4096), (16, 1024)):
    r = bench.plan_benchmark(wl, 20, 32768, wave, "gpu", trees=trees)
    r.pop("root_stats")
    print(f"trees {trees} wave {wave}:", r, flush=True)

# latency of a Plan() at the reference's default budget (2000 iterations, one tree)
import time  # noqa: E402
for wave in (0, 64, 256, 1024):
    r = bench.plan_benchmark(wl, 20, 2000, wave, "gpu", trees=1)
    r.pop("root_stats")
    print(f"default budget, wave {wave}: plan {r['plan_ms']:.2f} ms, search {r['search_ms']:.2f} ms", flush=True)
r = bench.plan_benchmark(wl, 20, 2000, 1, "cpu", trees=1)
print(f"default budget, CPU oracle sequential: plan {r['plan_ms']:.2f} ms", flush=True)
