#!/usr/bin/env python
"""Plan() time split (host select / fill / engine / backup) at several wave sizes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["BMCTS_DEBUG"] = "1"
import bench  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2_merge_sbg"
for wave in (256, 1024, 4096):
    r = bench.plan_benchmark(wl, 20, 32768, wave, "gpu")
    r.pop("root_stats")
    print(wave, r, flush=True)
