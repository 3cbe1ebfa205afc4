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
