import os, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
os.environ["BMCTS_DEBUG"] = "1"
import bench
for trees, threads in ((1, 1), (32, 16)):
    r = bench.plan_benchmark("c5_rsbg", 20, 32768, 0, "gpu", trees=trees, threads=threads)
    r.pop("root_stats")
    print("trees", trees, "it/s %.4g" % r["iterations_per_s"], "plan_ms %.1f" % r["plan_ms"], flush=True)
