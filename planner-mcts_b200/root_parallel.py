"""Root-parallel search over ranks (one process per GPU).

The reference's root parallelism is `Mcts::NumParallelMcts` trees, each reseeded
(`choose_random_seed`, mcts_state/mcts_state_hypothesis.cpp:97-101), whose root statistics are
merged (`merge_node_statistics`, mcts_statistics/mcts_neural_cost_constrained_statistic.hpp:157-163).
Here every rank owns one GPU and one tree (tree index = rank -> distinct Philox stream); the
only exchange is one all-reduce (sum, float64) of the flattened root statistics -- a few hundred
bytes over NVLink with the `nccl` backend, or over `gloo` in the CPU tests -- after which every
rank takes the same decision from the merged statistics.
"""
import numpy as np
import torch
import torch.distributed as dist


def merge_root_stats(stats, device=None, group=None):
    """all-reduce (sum) of a float64 root-statistics vector; identity without a process group"""
    stats = np.ascontiguousarray(stats, dtype=np.float64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats
    t = torch.from_numpy(stats.copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def root_parallel_plan(planner, delta_time, observed_world, device=None, group=None):
    """Plan() on this rank's tree, merge the root statistics over all ranks, decide from the
    merged statistics.  Returns (best_action, merged_stats); identical on every rank."""
    rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    planner.set_tree_index(rank)
    planner.Plan(delta_time, observed_world)
    merged = merge_root_stats(planner.root_stats(), device=device, group=group)
    best = planner.decide_from_root_stats(merged)
    return best, merged
