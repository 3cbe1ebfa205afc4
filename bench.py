#!/usr/bin/env python
"""bench.py -- rollout agent-steps/s of the MCTS leaf-evaluation path on B200.

Contract: `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line.
A step = one rollout batch (bmcts_rollout_batch) over `--rollouts` synthetic leaves of the
configured BASELINE.json workload -- default: configs[4], the largest one ("RSBG
risk-constrained, 32 agents, 64 hypotheses"), horizon 20 steps with terminal early exit.  The
leaves come in the unique-state form of the C ABI (a table of rollouts / 16 distinct leaf states,
each rolled out under 16 freshly sampled hypothesis assignments: SURVEY.md Appendix G).
`value` is measured with the inputs resident in HBM (CUDA events on the engine's stream); `e2e`
is the same metric through the C-ABI call with pinned HOST buffers (host->device copies of the
inputs and the device->host path of the results inside the timed region).  The same line carries
`workloads` (the other four BASELINE.json configs, short runs), `likelihood` (the belief-update
kernel at C5 size), `plan` (MCTS iterations/s of one Plan() with 10^6 iterations on C5,
root-parallel over the ranks) and `plan_default_budget` (the reference's default budget: 2000
iterations, one tree).  `--impl reference` times the CPU oracle (the reference itself cannot be
built here) on all host cores on a bounded sample of the same workload.
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


def hbm_bytes_per_rollout(A, share=1):
    """algorithmic HBM bytes of one rollout: its share of a leaf state (16 B per agent + alive
    mask; `share` rollouts start from the same state), one hypothesis id per agent, the state
    index, and the 32-byte result record"""
    return (A * 16 + 4) / share + A + (4 if share > 1 else 0) + 32


def input_bytes_per_step(A, n, share):
    ns = n // share if share > 1 else n
    return int(ns * (A * 16 + 4) + n * A + (4 * n if share > 1 else 0))


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
              b + "Mcts::Engine::MaxThreads": threads,
              # the workload's Euler step of the ego's macro action (10 per 0.5 s)
              b + "EgoBehavior::BehaviorMotionPrimitives::IntegrationTimeDelta": float(cfg.ego_integration_dt)}
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


def make_batch(scn, meta, n, share, seed):
    """n synthetic leaves in the unique-state form: n / share distinct leaf states (poses around
    the workload's placement), each rolled out `share` times under independently sampled
    hypothesis ids; share == 1 is the plain form (one state per rollout)."""
    from planner_mcts_b200 import scenarios
    ns = n // share if share > 1 else n
    pose, alive, _ = scenarios.leaves_for(scn, meta, ns, seed=seed)
    rng = np.random.default_rng(seed + 7919)
    hyp = rng.integers(0, max(1, scn.n_hypotheses), size=(scn.n_agents, n), dtype=np.uint8)
    index = None
    if share > 1:
        # the rollouts of one state are adjacent, the way a search emits them (node by node, each
        # under its freshly sampled hypotheses); the C ABI takes the state table state-major
        index = np.repeat(np.arange(ns, dtype=np.uint32), share)
        pose = np.ascontiguousarray(pose.transpose(1, 0, 2))
    return pose, alive, hyp, index


class Arm:
    """device-resident and host (pinned) copies of one batch + the calls that roll it out"""

    def __init__(self, eng, scn, meta, n, share, seed, dev, rank, want_host=True):
        import torch
        from planner_mcts_b200 import _abi
        self.eng, self.n, self.rank, self._abi, self.torch = eng, n, rank, _abi, torch
        pose, alive, hyp, index = make_batch(scn, meta, n, share, seed)
        self.indexed = index is not None
        self.ns = pose.shape[0] if self.indexed else pose.shape[1]
        t = lambda a: None if a is None else torch.from_numpy(a)
        host = [t(pose), t(alive.view(np.int32)), t(hyp), t(None if index is None else index.view(np.int32))]
        self.dev = [None if x is None else x.to(dev) for x in host]
        self.d_res = torch.zeros(n, 8, dtype=torch.float32, device=dev)
        if want_host:
            self.host = [None if x is None else x.pin_memory() for x in host]
            self.h_res = torch.zeros(n, 8, dtype=torch.float32).pin_memory()
        torch.cuda.synchronize()

    def _call(self, mem, bufs, res, i):
        pose, alive, hyp, index = bufs
        self.eng.rollout_raw(self.n, mem, pose, alive, hyp, None, res, seed=1000, stream=self.rank,
                             first_rollout_id=i * self.n, n_states=self.ns if self.indexed else 0,
                             state_index=index)

    def step_device(self, i):
        self._call(self._abi.MEM_DEVICE, self.dev, self.d_res, i)

    def step_host(self, i):   # returns when the results are in the pinned host buffer
        self._call(self._abi.MEM_HOST, self.host, self.h_res, i)


def measure(eng, arm, steps, warmup, stream, barrier, sampler_index=None, want_e2e=True):
    """K timed steps device-resident (CUDA events on the engine's stream), the agent-steps they
    executed (recounted outside the timed region: rollouts are deterministic per (leaf, id)), the
    per-launch kernel time, and the same steps end to end from pinned host buffers."""
    import torch
    from planner_mcts_b200 import RESULT_DTYPE
    for i in range(warmup):
        arm.step_device(i)
    barrier()
    launches0 = eng.launch_count()
    sampler = ClockSampler(sampler_index) if sampler_index is not None else None
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(steps):
        arm.step_device(warmup + i)
    ev1.record(stream)
    barrier()
    out = {"elapsed_ms": ev0.elapsed_time(ev1), "launches": eng.launch_count() - launches0,
           "clocks": sampler.summary() if sampler else None}
    kernel_ms, total = [], 0
    count_steps = min(steps, 4)   # the step totals of the other steps differ by < 0.1 %
    for i in range(count_steps):
        arm.step_device(warmup + i)
        eng.sync()
        kernel_ms.append(eng.last_kernel_ms())
        res = arm.d_res.cpu().numpy().view(RESULT_DTYPE).reshape(-1)
        total += int(res["agent_steps"].sum())
    out["agent_steps_per_step"] = total / count_steps
    out["kernel_ms"] = float(np.mean(kernel_ms))
    out["mean_rollout_steps"] = float(res["n_steps"].mean())
    out["lane_T"] = None
    if want_e2e:
        for i in range(max(1, min(warmup, 2))):
            arm.step_host(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            arm.step_host(warmup + i)
        torch.cuda.synchronize()
        out["e2e_s"] = time.perf_counter() - t0
        arm.step_host(warmup + count_steps - 1)
        last = arm.h_res.numpy().view(RESULT_DTYPE).reshape(-1)
        assert np.array_equal(last, res), "host-memory path and device-resident path disagree"
    return out


def roofline_of(m, A, coop, sm_mhz, peaks):
    flop = flop_per_agent_step(A, coop)
    achieved = m["agent_steps_per_step"] * flop / (m["kernel_ms"] * 1e-3) / 1e12
    peak = 2 * 128 * 148 * sm_mhz * 1e6 / 1e12
    return achieved, peak, flop


def likelihood_benchmark(device, peaks, sm_mhz):
    """bmcts_likelihood_batch at C5 size (32 agents x 64 hypotheses x 10^4 samples per world,
    hypothesis/idm/hypothesis_idm.cpp:40-126), 16 worlds per call, device-resident"""
    import torch
    from planner_mcts_b200 import RolloutEngine, _abi, scenarios
    cfg, scn, meta = scenarios.workload("c5_rsbg")
    worlds, samples = 16, 10000
    pose, alive, _ = scenarios.leaves_for(scn, meta, worlds, seed=11)
    rng = np.random.default_rng(0)
    action = rng.uniform(-3, 2, size=(scn.n_agents, worlds)).astype(np.float32)
    eng = RolloutEngine(cfg, scn, device=device)
    dev = torch.device("cuda", device)
    d_pose = torch.from_numpy(pose).to(dev)
    d_alive = torch.from_numpy(alive.view(np.int32)).to(dev)
    d_action = torch.from_numpy(action).to(dev)
    d_prob = torch.zeros(scn.n_agents * scn.n_hypotheses * worlds, dtype=torch.float32, device=dev)
    ms = []
    for i in range(13):
        eng.likelihood_raw(worlds, _abi.MEM_DEVICE, d_pose, d_alive, d_action, d_prob,
                           n_samples=samples, n_buckets=100, seed=3 + i)
        eng.sync()
        if i >= 3:
            ms.append(eng.last_kernel_ms())
    k_ms = float(np.median(ms))
    evals = worlds * scn.n_agents * scn.n_hypotheses * samples
    # per evaluation: the IDM acceleration (F_idm = 48 FLOP of SURVEY.md 8d) -- the Philox draw and
    # the bucket test are integer work on top of it
    achieved = evals * 48.0 / (k_ms * 1e-3) / 1e12
    peak = 2 * 128 * 148 * sm_mhz * 1e6 / 1e12
    # host call on one world: what a Plan() pays for its belief update
    p1, a1, _ = scenarios.leaves_for(scn, meta, 1, seed=12)
    act1 = action[:, :1].copy()
    eng.likelihood(p1, a1, act1, n_samples=samples, n_buckets=100)
    t0 = time.perf_counter()
    for _ in range(10):
        eng.likelihood(p1, a1, act1, n_samples=samples, n_buckets=100)
    call_ms = (time.perf_counter() - t0) / 10 * 1e3
    return {"kernel": "bmcts::likelihood_kernel", "workload": "c5_rsbg", "worlds_per_call": worlds,
            "evaluations_per_call": evals, "kernel_ms": k_ms,
            "evaluations_per_s": evals / (k_ms * 1e-3),
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "flop_per_evaluation": 48},
            "belief_update_of_one_plan_ms": call_ms}


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
    pose, alive, hyp, index = make_batch(scn, meta, n, args.share, seed=1)
    if index is not None:
        pose = np.ascontiguousarray(pose.transpose(1, 0, 2))   # tests/oracle.py takes [A, n_states, 4]
    oracle.load()
    kw = dict(state_index=index, n_threads=cores)
    w = max(64, n // 16)
    for _ in range(args.warmup):
        oracle.rollout(cfg, scn, pose, alive, hyp[:, :w], **dict(kw, state_index=None if index is None else index[:w]))
    total, t0 = 0, time.perf_counter()
    for s in range(args.steps):
        r = oracle.rollout(cfg, scn, pose, alive, hyp, first_rollout_id=s * n, **kw)
        total += int(r["result"]["agent_steps"].sum())
    dt = time.perf_counter() - t0
    val = total / dt
    sample = f"{n} rollouts/step x {args.steps} steps of {args.workload} (horizon {args.horizon})"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args, meta),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "plan": None if args.no_plan else {k: v for k, v in plan_benchmark(
            args.workload, args.horizon, args.cpu_plan_iterations, 1, "cpu", trees=cores).items()
            if k != "root_stats"},
        "note": "reference cannot be compiled here (Bazel + BARK + mamcts absent); this is the "
                "in-repo CPU oracle, which has none of BARK's heap cloning and is therefore "
                "faster than the real reference would be; the sample is a bounded batch of the "
                "same workload (the GPU arm's step is rollouts_per_step_per_gpu of them)",
    }
    print(json.dumps(line), flush=True)


def workload_config(args, meta):
    """`config` of the JSON line: the SAME on both arms for the same command line (the reference
    arm's step is a bounded sample of this workload, its size is in cpu_baseline.sample; what a
    run measured on top -- rollout lengths, rates, the timed region -- goes to `run`)"""
    in_bytes = input_bytes_per_step(meta["A"], args.rollouts, args.share)
    return {"workload": args.workload, "agents": meta["A"], "hypotheses": meta["H"],
            "horizon": args.horizon, "ego_substeps": 10,
            "leaf_form": ("unique-state table: rollouts / %d distinct leaf states, each under %d sampled "
                          "hypothesis assignments" % (args.share, args.share)) if args.share > 1
            else "one state per rollout",
            "rollouts_per_step_per_gpu": args.rollouts,
            "l2": "inputs larger than L2 (%.0f MB of leaf states, hypothesis ids and indices per "
                  "step)" % (in_bytes / 1e6)}


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


# rollouts per step of the short runs: enough rollouts per lane (>= 35) that the tail of the
# persistent grid is below 2 %
OTHER_WORKLOADS = [("c1_highway_mdp", 1 << 22), ("c2_merge_sbg", 1 << 22), ("c3_merge_risk", 1 << 21),
                   ("c4_coop", 1 << 21), ("c5_rsbg", 1 << 21)]


def run_ours(args):
    import torch
    import torch.distributed as dist

    from planner_mcts_b200 import RolloutEngine, _abi, scenarios

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

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks, peak_kind = measured_peaks()

    # ---- the headline workload
    cfg, scn, meta = scenarios.workload(args.workload, horizon=args.horizon)
    A, n = meta["A"], args.rollouts
    coop = cfg.state_kind == _abi.STATE_COOPERATIVE
    eng = RolloutEngine(cfg, scn, device=local_rank)
    stream = torch.cuda.ExternalStream(eng.stream_ptr(), device=dev)
    arm = Arm(eng, scn, meta, n, args.share, 1 + rank, dev, rank)
    m = measure(eng, arm, args.steps, args.warmup, stream, barrier, sampler_index=local_rank)
    clocks = m["clocks"]
    sm_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    total_steps = m["agent_steps_per_step"] * args.steps
    del arm
    torch.cuda.empty_cache()

    # ---- the other BASELINE.json configs (short runs, device-resident + end to end)
    table = {}
    if not args.no_workloads:
        for name, n_w in OTHER_WORKLOADS:
            if name == args.workload:
                continue
            for tag, ego_dt, hz in ((name, 0.05, args.horizon),) + (
                    (("c2_merge_sbg@reference_ego_dt_0.02", 0.02, args.horizon),
                     ("c2_merge_sbg@rollout_cap_1000", 0.05, 1000)) if name == "c2_merge_sbg" else ()):
                cfg_w, scn_w, meta_w = scenarios.workload(name, horizon=hz, ego_dt=ego_dt)
                n_t = n_w
                eng_w = RolloutEngine(cfg_w, scn_w, device=local_rank)
                st_w = torch.cuda.ExternalStream(eng_w.stream_ptr(), device=dev)
                arm_w = Arm(eng_w, scn_w, meta_w, n_t, args.share, 11 + rank, dev, rank)
                mw = measure(eng_w, arm_w, 5, 3, st_w, barrier)
                coop_w = cfg_w.state_kind == _abi.STATE_COOPERATIVE
                ach, peak, flop = roofline_of(mw, meta_w["A"], coop_w, sm_mhz, peaks)
                table[tag] = {
                    "agents": meta_w["A"], "hypotheses": meta_w["H"], "horizon": hz,
                    "rollouts_per_step": n_t, "ego_substeps": 25 if ego_dt == 0.02 else 10,
                    "value": mw["agent_steps_per_step"] * 5 / (mw["elapsed_ms"] * 1e-3),
                    "e2e": mw["agent_steps_per_step"] * 5 / mw["e2e_s"],
                    "kernel_ms": mw["kernel_ms"], "mean_rollout_steps": mw["mean_rollout_steps"],
                    "roofline_frac": ach / peak, "flop_per_agent_step": flop,
                    "kernel": "bmcts::rollout_lane_coop_kernel" if coop_w else "bmcts::rollout_lane_kernel"}
                del arm_w, eng_w
                torch.cuda.empty_cache()

    # ---- belief-update kernel
    lik = None
    if rank == 0 and not args.no_workloads:
        lik = likelihood_benchmark(local_rank, peaks, sm_mhz)

    # ---- Plan() arm: one BehaviorUCT*::Plan() of --plan-iterations MCTS iterations in total,
    # root-parallel over the ranks (tree indices are disjoint), root statistics merged with one
    # NCCL all-reduce (BASELINE.json configs[4]: 10^6 rollouts per Plan())
    plan, plan_all, plan_w16 = None, None, None
    if not args.no_plan:
        from planner_mcts_b200 import root_parallel
        cores = os.cpu_count() or 1
        # ncclAllReduce behind the C ABI (one rank: nothing to merge, and NCCL stays unloaded)
        comm = root_parallel.RootComm(local_rank) if world > 1 else None
        if comm is not None:
            comm.allreduce(np.zeros(8))   # connection set-up of the new communicator, once

        def plan_arm(threads, wave=None, trees_per_thread=2):
            trees = args.plan_trees if args.plan_trees > 0 else trees_per_thread * threads
            per_tree = max(1, args.plan_iterations // (world * trees))
            g = plan_benchmark(args.workload, args.horizon, per_tree,
                               args.plan_wave if wave is None else wave, "gpu",
                               device=local_rank, tree_index=rank, trees=trees, threads=threads)
            ms = torch.tensor([g["plan_ms"]], dtype=torch.float64, device=dev)
            barrier()   # plan_ms is the MAX over ranks: waiting for the slowest rank is in it already
            t_merge = time.perf_counter()
            st = np.asarray(g.pop("root_stats"), dtype=np.float64)
            if comm is not None:
                st = comm.allreduce(st)
            merge_ms = (time.perf_counter() - t_merge) * 1e3
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            n_act = (len(st) - 2) // 4
            cnt, ret = st[:n_act], st[n_act:2 * n_act]
            mean = np.where(cnt > 0, ret / np.maximum(cnt, 1), -np.inf)
            return dict(g, iterations=int(st[-2]), plan_ms=float(ms[0]), merge_ms=merge_ms,
                        # one Plan() = the ranks' searches + the merge of their root statistics
                        iterations_per_s=float(st[-2]) / ((float(ms[0]) + merge_ms) * 1e-3),
                        merged_best_action=int(np.argmax(mean)), trees=world * trees,
                        workload=args.workload, host_cores=cores, host_threads_per_rank=threads,
                        merge="bmcts_allreduce_root (ncclAllReduce behind the C ABI)")

        # The scaling series: every rank searches with ITS SHARE of the host -- cores / 8 threads,
        # the share of one GPU on an 8-GPU box -- whatever N is, so that N ranks add N shares and
        # the series shows whether the path itself scales.  (Plan() is bound by the host tree
        # search: with cores / N threads per rank the N = 1 run would already use every core and
        # more GPUs could add nothing.)
        share = max(1, cores // 8)
        plan = plan_arm(share)
        plan["threads_note"] = "host threads per rank = cores / 8 at every N (per-GPU share of the host)"
        if world == 1:
            plan_all = plan_arm(max(1, min(16, cores)))   # one rank with (up to 16 of) all cores
            # the same budget in waves of 16 leaves -- the largest wave whose root estimates stay
            # within Monte-Carlo noise of the sequential search -- with four
            # trees per thread in flight to cover the latency of such a wave
            plan_w16 = plan_arm(max(1, min(16, cores)), wave=16, trees_per_thread=4)
        if comm is not None:
            comm.close()

    # ---- reduce over ranks: time = max, work = sum
    t = torch.tensor([m["elapsed_ms"], m["e2e_s"] * 1e3], dtype=torch.float64, device=dev)
    w = torch.tensor([float(total_steps), float(m["launches"])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    elapsed_ms, e2e_ms = float(t[0]), float(t[1])
    total_all, launches_all = float(w[0]), int(w[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = total_all / (elapsed_ms * 1e-3)
    e2e_value = total_all / (e2e_ms * 1e-3)
    k_ms = m["kernel_ms"]
    achieved_tflops, peak_tflops, flop = roofline_of(m, A, coop, sm_mhz, peaks)
    peak_tflops_max = 2 * 128 * 148 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
    alg_bytes = n * hbm_bytes_per_rollout(A, args.share)
    hbm_gbs = alg_bytes / (k_ms * 1e-3) / 1e9
    in_bytes = input_bytes_per_step(A, n, args.share)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, meta),
        "run": {"mean_rollout_steps": m["mean_rollout_steps"],
                "rollouts_per_s": n * world / (elapsed_ms / args.steps * 1e-3),
                "timed_region_s": elapsed_ms * 1e-3, "cpus_bound_per_rank": numa},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes,
                "d2h_bytes_per_step": int(n * 32)},
        "gpu_launches": launches_all,
        "roofline": {"bound": "fp32", "achieved": achieved_tflops, "peak": peak_tflops,
                     "unit": "TFLOP/s", "frac": achieved_tflops / peak_tflops,
                     "peak_at_max_clock": peak_tflops_max,
                     "peak_def": "2 x 128 FP32 lanes x 148 SMs x SM clock under load (%s MHz); "
                                 "algorithmic FLOPs per agent-step = %.0f (SURVEY.md 8d)" %
                                 (sm_mhz, flop),
                     "kernel": "bmcts::rollout_lane_coop_kernel" if coop else "bmcts::rollout_lane_kernel",
                     "kernel_ms": k_ms, "traffic": measured_traffic(args.workload, n),
                     "traffic_algorithmic": alg_bytes},
        "roofline_hbm": {"bound": "hbm", "achieved": hbm_gbs, "peak": peaks.get("hbm_gbs"),
                         "unit": "GB/s", "frac": hbm_gbs / peaks.get("hbm_gbs", 6650.0),
                         "peak_kind": peak_kind, "traffic": None},
    }
    if table:
        line["workloads"] = table
    if lik is not None:
        line["likelihood"] = lik
    if plan is not None:
        line["plan"] = plan
    if plan_all is not None:
        line["plan_all_host_threads"] = plan_all
    if plan_w16 is not None:
        line["plan_wave16"] = plan_w16
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, cfg, scn, meta)
        if plan is not None:
            cores = os.cpu_count() or 1
            c = plan_benchmark(args.workload, args.horizon, args.cpu_plan_iterations, 1, "cpu")
            c.pop("root_stats")
            line["cpu_baseline"]["plan"] = c   # the reference's default: one sequential tree
            c = plan_benchmark(args.workload, args.horizon, args.cpu_plan_iterations, 1, "cpu", trees=cores)
            c.pop("root_stats")
            line["cpu_baseline"]["plan_all_cores"] = c   # NumParallelMcts = cores, UseMultiThreading
    if world == 1 and not args.no_plan:
        # the reference's default budget (Mcts::MaxNumIterations 2000, one tree, default wave) on
        # c2 with 20-step rollouts: what one BARK agent pays per Plan()
        gd = plan_benchmark("c2_merge_sbg", 20, 2000, 0, "gpu", device=local_rank)
        cd = plan_benchmark("c2_merge_sbg", 20, 2000, 1, "cpu")
        line["plan_default_budget"] = {
            "workload": "c2_merge_sbg", "iterations": 2000, "trees": 1, "wave": "default",
            "gpu_plan_ms": gd["plan_ms"], "gpu_plan_ms_all": gd["plan_ms_all"],
            "cpu_plan_ms": cd["plan_ms"], "speedup": cd["plan_ms"] / gd["plan_ms"]}
        if not args.no_workloads:
            # the same budget on the other workloads (the CPU search pays per agent and step,
            # the GPU wave per pass of its longest rollout)
            others = {}
            for name in ("c1_highway_mdp", "c3_merge_risk", "c4_coop", "c5_rsbg"):
                g_ = plan_benchmark(name, 20, 2000, 0, "gpu", device=local_rank)
                c_ = plan_benchmark(name, 20, 2000, 1, "cpu")
                others[name] = {"gpu_plan_ms": g_["plan_ms"], "cpu_plan_ms": c_["plan_ms"],
                                "speedup": c_["plan_ms"] / g_["plan_ms"]}
            line["plan_default_budget"]["other_workloads"] = others
            # longer rollout caps on c2 (the reference's default cap is 1000 steps): a wave lasts
            # as long as its longest rollout, so this is where the wave-batched search gains least
            caps = {}
            for hz in (100, 1000):
                g_ = plan_benchmark("c2_merge_sbg", hz, 2000, 0, "gpu", device=local_rank)
                c_ = plan_benchmark("c2_merge_sbg", hz, 2000, 1, "cpu")
                caps[str(hz)] = {"gpu_plan_ms": g_["plan_ms"], "cpu_plan_ms": c_["plan_ms"],
                                 "speedup": c_["plan_ms"] / g_["plan_ms"]}
            line["plan_default_budget"]["rollout_cap_steps"] = caps
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args, cfg, scn, meta):
    """oracle ('port') on the host cores, bounded to roughly 10-20 s of CPU work; the same
    per-call batch as the reference arm (--cpu-rollouts)"""
    import oracle
    cores = os.cpu_count() or 1
    n = args.cpu_rollouts
    pose, alive, hyp, index = make_batch(scn, meta, n, args.share, seed=1)
    if index is not None:
        pose = np.ascontiguousarray(pose.transpose(1, 0, 2))   # tests/oracle.py takes [A, n_states, 4]
    oracle.load()
    w = max(64, n // 64)
    oracle.rollout(cfg, scn, pose, alive, hyp[:, :w], n_threads=cores,
                   state_index=None if index is None else index[:w])  # warm-up

    def timed(n_threads, budget_s, m):
        steps, calls, t0 = 0, 0, time.perf_counter()
        while time.perf_counter() - t0 < budget_s:
            r = oracle.rollout(cfg, scn, pose, alive, hyp[:, :m], first_rollout_id=calls * n,
                               n_threads=n_threads,
                               state_index=None if index is None else index[:m])
            steps += int(r["result"]["agent_steps"].sum())
            calls += 1
        dt = time.perf_counter() - t0
        return steps / dt, calls, dt

    m1 = max(256, n // 16)
    one_thread, c1, t1 = timed(1, 4.0, m1)
    val, c2, t2 = timed(cores, 10.0, n)
    return {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
            "one_thread_value": one_thread,
            "sample": f"{c2 * n} rollouts of {args.workload} (horizon {args.horizon}) on {cores} "
                      f"threads in {t2:.1f} s ({n} per call); {c1 * m1} rollouts on 1 thread in {t1:.1f} s"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5_rsbg")
    ap.add_argument("--horizon", type=int, default=20)
    ap.add_argument("--rollouts", type=int, default=1 << 22, help="rollouts per step per GPU")
    ap.add_argument("--share", type=int, default=16,
                    help="rollouts per distinct leaf state (unique-state form of the C ABI); 1 = one "
                         "state per rollout")
    ap.add_argument("--cpu-rollouts", type=int, default=65536,
                    help="CPU arms (reference arm, cpu_baseline): rollouts per call")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-plan", action="store_true", help="skip the Plan() iterations/s arms")
    ap.add_argument("--no-workloads", action="store_true",
                    help="skip the table of the other workloads and the likelihood kernel")
    ap.add_argument("--no-numa-bind", action="store_true",
                    help="multi-rank runs: do not bind the process to the CPUs next to its GPU")
    ap.add_argument("--plan-iterations", type=int, default=1 << 20,
                    help="MCTS iterations of the Plan() arm in total, split over ranks and trees "
                         "(BASELINE.json configs[4]: 10^6 rollouts per Plan())")
    ap.add_argument("--plan-wave", type=int, default=0,
                    help="iterations per engine batch; 0 = the planner's default (iterations / 128, "
                         "at most 256), which keeps root values within Monte-Carlo noise of the "
                         "sequential search")
    ap.add_argument("--plan-trees", type=int, default=0,
                    help="root-parallel trees per process (Mcts::NumParallelMcts); 0 = two per host "
                         "thread, with host threads = cores / ranks, at most 16")
    ap.add_argument("--cpu-plan-iterations", type=int, default=2000)
    args = ap.parse_args()
    if args.share < 1 or args.rollouts % args.share or args.cpu_rollouts % args.share:
        raise SystemExit("--rollouts and --cpu-rollouts must be multiples of --share")
    if args.impl == "ours":
        args.warmup = max(args.warmup, 3)   # timing rule: at least three untimed warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
