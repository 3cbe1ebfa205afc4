"""Root-parallel search over ranks: world_size-2 `gloo` run on CPU (the N>1 path of the
planner; on the GPU box the same code runs over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
import planner_worlds as W
from planner_mcts_b200 import planner as P
from planner_mcts_b200 import root_parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        world = W.observed_world(1, 2.0, 5.0, 4.0, goal_center=(150.0, -1.75))
        pl = oracle.make_planner(P.BehaviorUCTHypothesis, W.base_params(400), W.hypotheses())
        best, merged = root_parallel.root_parallel_plan(pl, 0.2, world)
        out[rank] = (best, merged.tolist(), pl.last_num_iterations)
    finally:
        dist.destroy_process_group()


def test_two_ranks_merge_root_statistics_over_gloo():
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        res = dict(out)
    assert set(res) == {0, 1}
    (b0, s0, it0), (b1, s1, it1) = res[0], res[1]
    assert b0 == b1                       # every rank takes the same decision
    assert s0 == s1                       # from bit-identical merged statistics
    assert it0 == it1 == 800              # iterations add up over the trees
    # the merge equals the sum of the two trees run in one process
    world = W.observed_world(1, 2.0, 5.0, 4.0, goal_center=(150.0, -1.75))
    total = None
    for tree in range(2):
        pl = oracle.make_planner(P.BehaviorUCTHypothesis, W.base_params(400), W.hypotheses())
        pl.set_tree_index(tree)
        pl.Plan(0.2, world)
        st = pl.root_stats()
        total = st if total is None else total + st
    assert np.array_equal(np.array(s0), total)


def test_merge_is_identity_without_process_group():
    st = np.arange(10, dtype=np.float64)
    assert np.array_equal(root_parallel.merge_root_stats(st), st)


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_collective_through_the_c_abi_alone():
    """tests/cpp/test_comm.cpp: bmcts_comm_* / bmcts_allreduce_root /
    bmcts_planner_allreduce_root_stats from a C++ program that links libbmcts.so and nothing else
    (no Python, no torch); one rank per visible GPU, at most two"""
    import subprocess
    from planner_mcts_b200 import _abi
    _abi.load_library()
    build = os.path.join(ROOT, "tests", "cpp", "_build")
    os.makedirs(build, exist_ok=True)
    exe = os.path.join(build, "test_comm")
    libdir = os.path.dirname(_abi.LIB_PATH)
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(cuda, "include"), os.path.join(ROOT, "tests", "cpp", "test_comm.cpp"),
                    "-o", exe, "-L", libdir, "-lbmcts", "-L", os.path.join(cuda, "lib64"), "-lcudart",
                    "-lpthread", f"-Wl,-rpath,{libdir}"], check=True, capture_output=True, text=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.strip().endswith("ok")


@pytest.mark.gpu
def test_root_comm_single_rank_and_planner_merge():
    """RootComm (the ctypes view of bmcts_comm_*) with one rank: the all-reduce is the identity
    and the planner's merged decision equals its own"""
    comm = root_parallel.RootComm(0, nranks=1, rank=0)
    st = np.arange(9, dtype=np.float64) * 0.25
    assert np.array_equal(comm.allreduce(st), st)
    world = W.observed_world(1, 2.0, 5.0, 4.0, goal_center=(150.0, -1.75))
    pl = P.BehaviorUCTHypothesis(W.base_params(300), W.hypotheses())
    traj = pl.Plan(0.2, world)
    best, stats = pl.last_macro_action, pl.root_stats()
    assert pl.allreduce_root_stats(comm) == best
    assert np.array_equal(pl.root_stats(), stats)
    assert np.array_equal(pl.last_trajectory, traj)
    comm.close()


@pytest.mark.gpu
def test_one_plan_over_several_gpus_in_one_process():
    """Mcts::Engine::Devices: the trees of ONE Plan() in ONE process run on the listed GPUs (engines
    are spread round robin) and are merged on the host: the statistics are those of the same
    trees on a single GPU.  (On a one-GPU box the list names device 0 twice.)"""
    import torch
    ndev = torch.cuda.device_count()
    devices = [0, 1 % ndev]
    world = W.observed_world(2, 4.0, 6.0, 3.0, goal_center=(150.0, -1.75))
    out = []
    for devs in (None, devices):
        params = W.base_params(200)
        params[W.P0 + "Mcts::NumParallelMcts"] = 6
        params[W.P0 + "Mcts::UseMultiThreading"] = 1
        params[W.P0 + "Mcts::Engine::MaxThreads"] = 3
        if devs:
            params[W.P0 + "Mcts::Engine::Devices"] = devs
        pl = P.BehaviorUCTHypothesis(params, W.hypotheses())
        pl.Plan(0.2, world)
        out.append((pl.root_stats(), pl.last_macro_action))
    assert np.array_equal(out[0][0], out[1][0]) and out[0][1] == out[1][1]
    assert out[0][0][-2] == 6 * 200
