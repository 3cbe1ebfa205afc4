"""Root-parallel search over ranks (one process per GPU).

The reference's root parallelism is `Mcts::NumParallelMcts` trees, each reseeded
(`choose_random_seed`, mcts_state/mcts_state_hypothesis.cpp:97-101), whose root statistics are
merged (`merge_node_statistics`, mcts_statistics/mcts_neural_cost_constrained_statistic.hpp:157-163).
Here every rank owns one GPU and one tree (tree index = rank -> distinct Philox stream); the
only exchange is one all-reduce (sum, float64) of the flattened root statistics -- a few hundred
bytes over NVLink with the `nccl` backend, or over `gloo` in the CPU tests -- after which every
rank takes the same decision from the merged statistics.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _abi


class RootComm:
    """The collective behind the C ABI (include/bmcts_c.h bmcts_comm_*): one ncclAllReduce(sum,
    float64) of the root statistics.  A C++ host creates it the same way; here torch.distributed
    only ships rank 0's 128-byte NCCL id to the other ranks."""

    def __init__(self, device, nranks=None, rank=None, unique_id=None, group=None):
        self._lib = _abi.load_library()
        if nranks is None:
            on = dist.is_available() and dist.is_initialized()
            nranks = dist.get_world_size(group) if on else 1
            rank = dist.get_rank(group) if on else 0
        if unique_id is None:
            buf = (C.c_ubyte * _abi.COMM_ID_BYTES)()
            if rank == 0:
                st = self._lib.bmcts_comm_unique_id(buf)
                if st != _abi.OK:
                    raise _abi.BmctsError(st, "bmcts_comm_unique_id")
            if nranks > 1:
                t = torch.tensor(list(buf), dtype=torch.uint8)
                if dist.get_backend(group) == "nccl":
                    t = t.to(torch.device("cuda", device))
                dist.broadcast(t, src=0, group=group)
                buf = (C.c_ubyte * _abi.COMM_ID_BYTES)(*t.cpu().tolist())
            unique_id = bytes(buf)
        self.nranks, self.rank = nranks, rank
        self.handle = C.c_void_p()
        idbuf = (C.c_ubyte * _abi.COMM_ID_BYTES).from_buffer_copy(unique_id)
        st = self._lib.bmcts_comm_create(int(device), int(nranks), int(rank), idbuf, C.byref(self.handle))
        if st != _abi.OK:
            raise _abi.BmctsError(st, "bmcts_comm_create")

    def allreduce(self, stats):
        stats = np.ascontiguousarray(stats, dtype=np.float64).copy()
        st = self._lib.bmcts_allreduce_root(self.handle, stats.ctypes.data_as(C.POINTER(C.c_double)),
                                            len(stats))
        if st != _abi.OK:
            raise _abi.BmctsError(st, self._lib.bmcts_comm_last_error(self.handle).decode())
        return stats

    def close(self):
        if self.handle:
            self._lib.bmcts_comm_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


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
