#!/usr/bin/env python
"""bench.py -- rollout agent-steps/s of the MCTS leaf-evaluation path on B200.

Contract: `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line.
A step = one rollout batch (bmcts_rollout_batch) over `--rollouts` synthetic leaf states of
the configured BASELINE.json workload (default: configs[1], "BehaviorUCTHypothesis SBG, 8
behavior hypotheses, merge scenario, 10 agents"), horizon 20 steps with terminal early
exit.  `value` is measured with the leaves resident in HBM (CUDA events on the engine's
stream); `e2e` is the same metric through the C-ABI call with pinned HOST buffers
(host->device copy of the leaves and device->host read of the results inside the timed
region).  `--impl reference` times the CPU oracle (the reference itself cannot be built
here) on all host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# one CUDA stream per root-parallel tree: use all hardware work queues (must be set before the
# CUDA context exists, i.e. before torch touches the device)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "rollout agent-steps/sec"
UNIT = "agent-steps/s"


def flop_per_agent_step(A, cooperative=False):
    """algorithmic FLOPs per agent-step (FMA = 2)."""
    if cooperative:
        return 400.0 + (A - 1) * 168.0 + 680.0
    return ((A - 1) * (480.0 + 60.0 * (A - 1)) + 400.0 + (A - 1) * 168.0 + 680.0) / A


def hbm_bytes_per_rollout(A):
    """algorithmic HBM bytes of one rollout: leaf state in, 32-byte result record out"""
    return A * (16 + 1) + 4 + 32


class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons of one GPU during the timed region (NVML; falls
    back to nvidia-smi when pynvml is missing)"""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        bits = [getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)]
        return [sm, self.max_mhz] + [bool(r & b) for b in bits]

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                              "--format=csv,noheader,nounits"], capture_output=True, text=True,
                             timeout=5).stdout.strip()
        c = [x.strip() for x in out.split(",")]
        return [float(c[0]), float(c[1])] + [x == "Active" for x in c[2:6]]

    def run(self):
        while not self.stop_flag.is_set():
            try:
                self.rows.append(self._sample_nvml() if self.nvml else self._sample_smi())
            except Exception:
                pass
            self.stop_flag.wait(0.02 if self.nvml else 0.15)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(r[0] for r in self.rows)
        reasons = [n for i, n in enumerate(self.NAMES) if any(r[2 + i] for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.rows[0][1], "reasons": reasons,
                "samples": len(self.rows)}


def measured_traffic(workload, n):
    """dram bytes (read + write) of one launch of the dominant kernel from the committed
    `ncu --set full` capture of this workload at this batch size (profiles/traffic.json)"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            for row in json.load(f):
                if row["workload"] == workload and row["rollouts"] == n:
                    return row["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def plan_benchmark(workload, horizon, iterations, wave, backend, device=0, tree_index=0, trees=1,
                   threads=0):
    """MCTS iterations/s of one BehaviorUCT*::Plan() on the workload's scenario (second half of
    BASELINE.json's metric).  backend 'gpu' = the product (host tree search, waves of leaves on
    the GPU); 'cpu' = the same tree search on the CPU oracle, sequential (wave 1) like the
    reference's mcts::Mcts::search."""
    from planner_mcts_b200 import _abi, planner as P, scenarios
    cfg, scn, meta = scenarios.workload(workload, horizon=horizon)
    pose, alive, _ = scenarios.leaves_for(scn, meta, 1, seed=3)
    A, H = meta["A"], meta["H"]
    agents = [P.Agent(id=1 + a, x=float(pose[a, 0, 0]), y=float(pose[a, 0, 1]),
                      theta=float(pose[a, 0, 2]), v=float(pose[a, 0, 3])) for a in range(A)]
    world = P.ObservedWorld(ego_id=1, agents=agents, map=scn)
    b = "BehaviorUctBase::"
    params = {b + "Mcts::MaxNumIterations": iterations, b + "Mcts::MaxSearchTime": 1e9,
              b + "Mcts::MaxNumNodes": max(2000, iterations), b + "Mcts::Engine::WaveSize": wave,
              b + "MaxNearestAgents": A - 1, b + "Mcts::RandomHeuristic::MaxNumIterations": horizon,
              b + "EgoBehavior::AccelerationInputs": [0.0, 1.0, 4.0, -1.0, -8.0],
              b + "Mcts::NumParallelMcts": trees, b + "Mcts::UseMultiThreading": 1 if trees > 1 else 0,
              b + "Mcts::Engine::MaxThreads": threads}
    if cfg.add_safe_dist:
        params[b + "Mcts::State::EvaluationParameters::AddSafeDist"] = 1
        params[b + "Mcts::State::SplitSafeDistCollision"] = 1
        params["BehaviorUctRiskConstraint::DefaultAvailableRisk"] = [0.1, 0.0]
    hyps = [P.HypothesisIDM(scn.hypotheses[h], num_samples=1000) for h in range(H)]
    if cfg.state_kind == _abi.STATE_COOPERATIVE:
        cls, hyp_arg = P.BehaviorUCTCooperative, None
    elif cfg.state_kind == _abi.STATE_RISK_CONSTRAINT:
        cls, hyp_arg = P.BehaviorUCTRiskConstraint, hyps
    else:
        cls, hyp_arg = P.BehaviorUCTHypothesis, hyps
    if backend == "cpu":
        import oracle
        pl = oracle.make_planner(cls, params, hyp_arg)
    else:
        pl = cls(params, hyp_arg, device=device) if hyp_arg is not None else cls(params, device=device)
    pl.set_tree_index(tree_index)
    pl.Plan(0.2, world)   # warm-up: buffers, first launches
    walls = []
    for _ in range(3 if backend == "gpu" else 1):   # median of three Plan() calls (host noise)
        t0 = time.perf_counter()
        pl.Plan(0.2, world)
        walls.append(time.perf_counter() - t0)
    wall = sorted(walls)[len(walls) // 2]
    return {"iterations_per_s": pl.last_num_iterations / wall, "iterations": pl.last_num_iterations,
            "plan_ms": wall * 1e3, "plan_ms_all": [round(w * 1e3, 2) for w in walls],
            "search_ms": pl.last_solution_time, "wave": wave,
            "trees_per_process": trees, "threads_per_process": threads or (os.cpu_count() or 1),
            "iterations_per_tree": iterations,
            "nodes": pl.last_num_nodes, "best_action": pl.last_macro_action,
            "root_stats": pl.root_stats().tolist()}


def run_reference(args):
    """CPU arm: the oracle (kind 'port') on all host cores, bounded sample per step."""
    import oracle
    from planner_mcts_b200 import scenarios
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, scn, meta = scenarios.workload(args.workload, horizon=args.horizon)
    cores = os.cpu_count() or 1
    n = args.cpu_rollouts
    pose, alive, hyp = scenarios.leaves_for(scn, meta, n, seed=1)
    oracle.load()
    for w in range(args.warmup):
        oracle.rollout(cfg, scn, pose[:, :max(64, n // 16)], alive[:max(64, n // 16)],
                       hyp[:, :max(64, n // 16)], n_threads=cores)
    total, t0 = 0, time.perf_counter()
    for s in range(args.steps):
        r = oracle.rollout(cfg, scn, pose, alive, hyp, first_rollout_id=s * n, n_threads=cores)
        total += int(r["result"]["agent_steps"].sum())
    dt = time.perf_counter() - t0
    val = total / dt
    sample = f"{n} rollouts/step x {args.steps} steps of {args.workload} (horizon {args.horizon})"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "agents": meta["A"], "hypotheses": meta["H"],
                   "horizon": args.horizon, "rollouts_per_step": n},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "plan": None if args.no_plan else {k: v for k, v in plan_benchmark(
            args.workload, args.horizon, args.cpu_plan_iterations, 1, "cpu").items() if k != "root_stats"},
        "note": "reference cannot be compiled here (Bazel + BARK + mamcts absent); this is the "
                "in-repo CPU oracle, which has none of BARK's heap cloning and is therefore "
                "faster than the real reference would be",
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(index):
    """Multi-rank runs: keep this process (its pinned staging memory is first touched by it, and
    its tree-search threads) on the CPUs next to its GPU, as NVML reports them; the ranks of one
    box otherwise stream their host batches across the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def run_ours(args):
    import torch
    import torch.distributed as dist

    from planner_mcts_b200 import RESULT_DTYPE, RolloutEngine, _abi, scenarios

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the rollout engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 and not args.no_numa_bind else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg, scn, meta = scenarios.workload(args.workload, horizon=args.horizon)
    A, n = meta["A"], args.rollouts
    coop = cfg.state_kind == _abi.STATE_COOPERATIVE
    pose, alive, hyp = scenarios.leaves_for(scn, meta, n, seed=1 + rank)
    eng = RolloutEngine(cfg, scn, device=local_rank)
    stream = torch.cuda.ExternalStream(eng.stream_ptr(), device=dev)

    # ---- device-resident arm
    d_pose = torch.from_numpy(pose).to(dev)
    d_alive = torch.from_numpy(alive.view(np.int32)).to(dev)
    d_hyp = torch.from_numpy(hyp).to(dev)
    d_res = torch.zeros(n, 8, dtype=torch.float32, device=dev)
    d_stats = torch.zeros(8, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    def step_device(i):
        eng.rollout_raw(n, _abi.MEM_DEVICE, d_pose, d_alive, d_hyp, None, d_res, seed=1000,
                        stream=rank, first_rollout_id=i * n)

    def merge_root_stats():
        # root-parallel merge: sum of the per-rank root statistics over NVLink
        with torch.cuda.stream(stream):
            d_stats[:4] = d_res[:, :4].sum(0, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(d_stats)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step_device(i)
        merge_root_stats()
    barrier()
    launches0 = eng.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, total_steps = [], 0
    ev0.record(stream)
    for i in range(args.steps):
        step_device(args.warmup + i)
        merge_root_stats()
    ev1.record(stream)
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0
    clocks = sampler.summary()
    # agent-steps are deterministic per (leaf, rollout id): recount them outside the timed region
    for i in range(args.steps):
        step_device(args.warmup + i)
        eng.sync()
        kernel_ms.append(eng.last_kernel_ms())
        res = d_res.cpu().numpy().view(RESULT_DTYPE).reshape(-1)
        total_steps += int(res["agent_steps"].sum())
    mean_rollout_len = float(res["n_steps"].mean())

    # ---- end-to-end arm: pinned host buffers through the C ABI, copies inside the timed region
    h_pose = torch.from_numpy(pose).pin_memory()
    h_alive = torch.from_numpy(alive.view(np.int32)).pin_memory()
    h_hyp = torch.from_numpy(hyp).pin_memory()
    h_res = torch.zeros(n, 8, dtype=torch.float32).pin_memory()

    def step_host(i):
        eng.rollout_raw(n, _abi.MEM_HOST, h_pose, h_alive, h_hyp, None, h_res, seed=1000,
                        stream=rank, first_rollout_id=i * n)

    for i in range(max(1, args.warmup)):
        step_host(i)
    barrier()
    # the timed steps use the rollout ids of the device-resident arm, so the agent-steps
    # they execute are the ones counted above (checked on the last step, outside the timing)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_host(args.warmup + i)   # returns when the results are in the pinned host buffer
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_steps = total_steps
    last = h_res.numpy().view(RESULT_DTYPE).reshape(-1)
    assert np.array_equal(last, res), "host-memory path and device-resident path disagree"

    # ---- Plan() arm: one BehaviorUCT*::Plan per rank (root-parallel: tree index = rank), root
    # statistics merged with an NCCL all-reduce
    plan = None
    if not args.no_plan:
        threads = max(1, min(16, (os.cpu_count() or 1) // world))
        if args.plan_trees <= 0:
            # two trees per host thread: a thread interleaves their waves, so the GPU works on
            # one tree while the thread selects / backs up the other
            args.plan_trees = 2 * threads
        g = plan_benchmark(args.workload, args.horizon, args.plan_iterations, args.plan_wave, "gpu",
                           device=local_rank, tree_index=rank, trees=args.plan_trees,
                           threads=threads)
        from planner_mcts_b200 import root_parallel
        ms = torch.tensor([g["plan_ms"]], dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        t_merge = time.perf_counter()
        st = root_parallel.merge_root_stats(g.pop("root_stats"), device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        torch.cuda.synchronize()
        merge_ms = (time.perf_counter() - t_merge) * 1e3
        n_act = (len(st) - 2) // 4
        cnt, ret = st[:n_act], st[n_act:2 * n_act]
        mean = np.where(cnt > 0, ret / np.maximum(cnt, 1), -np.inf)
        plan = dict(g, iterations=int(st[-2]), plan_ms=float(ms[0]), merge_ms=merge_ms,
                    iterations_per_s=float(st[-2]) / (float(ms[0]) * 1e-3),
                    merged_best_action=int(np.argmax(mean)), trees=world)

    # ---- reduce over ranks: time = max, work = sum
    t = torch.tensor([elapsed_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    w = torch.tensor([float(total_steps), float(e2e_steps), float(launches)], dtype=torch.float64,
                     device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    elapsed_ms, e2e_ms = float(t[0]), float(t[1])
    total_all, e2e_all, launches_all = float(w[0]), float(w[1]), int(w[2])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = total_all / (elapsed_ms * 1e-3)
    e2e_value = e2e_all / (e2e_ms * 1e-3)
    peaks, peak_kind = measured_peaks()
    k_ms = float(np.mean(kernel_ms))
    steps_per_launch = total_steps / args.steps
    flop = flop_per_agent_step(A, coop)
    achieved_tflops = steps_per_launch * flop / (k_ms * 1e-3) / 1e12
    sm_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    peak_tflops = 2 * 128 * 148 * sm_mhz * 1e6 / 1e12
    peak_tflops_max = 2 * 128 * 148 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
    hbm_gbs = n * hbm_bytes_per_rollout(A) / (k_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "agents": A, "hypotheses": meta["H"],
                   "horizon": args.horizon, "rollouts_per_step_per_gpu": n,
                   "mean_rollout_steps": mean_rollout_len,
                   "l2": "inputs larger than L2 (leaf batch %.0f MB per step)" %
                         (n * (A * 17 + 4) / 1e6),
                   "rollouts_per_s": n * world / (elapsed_ms / args.steps * 1e-3),
                   "cpus_bound_per_rank": numa},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": int(n * (A * 17 + 4)), "d2h_bytes_per_step": int(n * 32)},
        "gpu_launches": launches_all,
        "roofline": {"bound": "fp32", "achieved": achieved_tflops, "peak": peak_tflops,
                     "unit": "TFLOP/s", "frac": achieved_tflops / peak_tflops,
                     "peak_at_max_clock": peak_tflops_max,
                     "peak_def": "2 x 128 FP32 lanes x 148 SMs x SM clock under load (%s MHz); "
                                 "algorithmic FLOPs per agent-step = %.0f (SURVEY.md 8d)" %
                                 (sm_mhz, flop),
                     "kernel": "bmcts::rollout_lane_coop_kernel" if coop else "bmcts::rollout_lane_kernel", "kernel_ms": k_ms, "traffic": measured_traffic(args.workload, n),
                     "traffic_algorithmic": n * hbm_bytes_per_rollout(A)},
        "roofline_hbm": {"bound": "hbm", "achieved": hbm_gbs, "peak": peaks.get("hbm_gbs"),
                         "unit": "GB/s", "frac": hbm_gbs / peaks.get("hbm_gbs", 6650.0),
                         "peak_kind": peak_kind, "traffic": None},
    }
    if plan is not None:
        line["plan"] = plan
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, cfg, scn, meta)
        if plan is not None:
            c = plan_benchmark(args.workload, args.horizon, args.cpu_plan_iterations, 1, "cpu")
            c.pop("root_stats")
            line["cpu_baseline"]["plan"] = c   # the reference's default: one sequential tree
            cores = os.cpu_count() or 1
            c = plan_benchmark(args.workload, args.horizon, args.cpu_plan_iterations, 1, "cpu", trees=cores)
            c.pop("root_stats")
            line["cpu_baseline"]["plan_all_cores"] = c   # NumParallelMcts = cores, UseMultiThreading
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args, cfg, scn, meta):
    """oracle ('port') on the host cores, bounded to roughly 10-20 s of CPU work"""
    import oracle
    from planner_mcts_b200 import scenarios
    cores = os.cpu_count() or 1
    n = 16384
    pose, alive, hyp = scenarios.leaves_for(scn, meta, n, seed=1)
    oracle.load()
    oracle.rollout(cfg, scn, pose[:, :256], alive[:256], hyp[:, :256], n_threads=cores)  # warm-up

    def timed(n_threads, budget_s):
        steps, calls, t0 = 0, 0, time.perf_counter()
        while time.perf_counter() - t0 < budget_s:
            r = oracle.rollout(cfg, scn, pose, alive, hyp, first_rollout_id=calls * n,
                               n_threads=n_threads)
            steps += int(r["result"]["agent_steps"].sum())
            calls += 1
        dt = time.perf_counter() - t0
        return steps / dt, calls, dt

    one_thread, c1, t1 = timed(1, 4.0)
    val, c2, t2 = timed(cores, 10.0)
    return {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
            "one_thread_value": one_thread,
            "sample": f"{c2 * n} rollouts of {args.workload} (horizon {args.horizon}) on {cores} "
                      f"threads in {t2:.1f} s; {c1 * n} rollouts on 1 thread in {t1:.1f} s"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2_merge_sbg")
    ap.add_argument("--horizon", type=int, default=20)
    ap.add_argument("--rollouts", type=int, default=1 << 20, help="leaf states per step per GPU")
    ap.add_argument("--cpu-rollouts", type=int, default=65536, help="reference arm: leaves per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-plan", action="store_true", help="skip the Plan() iterations/s arm")
    ap.add_argument("--no-numa-bind", action="store_true",
                    help="multi-rank runs: do not bind the process to the CPUs next to its GPU")
    ap.add_argument("--plan-iterations", type=int, default=32768)
    ap.add_argument("--plan-wave", type=int, default=0,
                    help="iterations per engine batch; 0 = the planner's default (iterations / 128, "
                         "at most 256), which keeps root values within Monte-Carlo noise of the "
                         "sequential search")
    ap.add_argument("--plan-trees", type=int, default=0,
                    help="root-parallel trees per process (Mcts::NumParallelMcts); 0 = two per host "
                         "thread, with host threads = cores / ranks, at most 16")
    ap.add_argument("--cpu-plan-iterations", type=int, default=4000)
    args = ap.parse_args()
    if args.impl == "ours":
        args.warmup = max(args.warmup, 3)   # timing rule: at least three untimed warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
