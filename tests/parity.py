"""Comparison helpers: CUDA path vs CPU oracle."""
import numpy as np

INT_FIELDS = ("flags", "n_steps", "agent_steps", "final_alive")
FLOAT_FIELDS = ("ret_ego", "cost0", "cost1", "step_len")
# north_star tolerance for returns / costs / poses: 1e-4 relative (FP32).  The shared
# elementary math makes the match bit-exact in practice; the tests check both.
RTOL = 1e-4


def assert_results_match(gpu, cpu, exact=True):
    g, c = gpu["result"], cpu["result"]
    for f in INT_FIELDS:  # bit-exact: flags, terminal index, counts
        bad = np.nonzero(g[f] != c[f])[0]
        assert bad.size == 0, f"{f}: {bad.size} mismatches, first at {bad[:5]}: " \
                              f"gpu {g[f][bad[:5]]} cpu {c[f][bad[:5]]}"
    for f in FLOAT_FIELDS:
        np.testing.assert_allclose(g[f], c[f], rtol=RTOL, atol=1e-5, err_msg=f)
        if exact:
            assert np.array_equal(g[f].view(np.uint32), c[f].view(np.uint32)), f"{f} not bit-exact"
    for k in ("ret_agents", "final_pose"):
        if k in gpu and k in cpu:
            np.testing.assert_allclose(gpu[k], cpu[k], rtol=RTOL, atol=1e-4, err_msg=k)
            if exact:
                assert np.array_equal(gpu[k].view(np.uint32), cpu[k].view(np.uint32)), \
                    f"{k} not bit-exact"


def assert_steps_match(gpu, cpu, exact=True, expanded=None):
    """expanded: boolean mask of the batch elements that were expanded (elements submitted with
    BMCTS_ACTION_NONE have no step outputs: the buffers keep whatever they held)"""
    def sel(a):
        return a if expanded is None else a[..., expanded] if a.ndim < 3 else a[:, expanded]
    for k in ("flags", "next_alive"):
        assert np.array_equal(sel(gpu[k]), sel(cpu[k])), k
    for k in ("next_pose", "reward", "cost", "last_action"):
        g, c = np.ascontiguousarray(sel(gpu[k])), np.ascontiguousarray(sel(cpu[k]))
        np.testing.assert_allclose(g, c, rtol=RTOL, atol=1e-4, err_msg=k)
        if exact:
            assert np.array_equal(g.view(np.uint32), c.view(np.uint32)), f"{k} not bit-exact"
