"""The C++ planner classes (planner-mcts_b200/host/behavior_uct.hpp) used directly, the way
the reference's C++ tests use theirs: tests/cpp/test_cpp_api.cpp built against the oracle-backed
planner library."""
import os
import subprocess

import pytest

import oracle
from conftest import has_gpu


def test_cpp_planner_api():
    oracle.build()
    exe = os.path.join(oracle.ORACLE_DIR, "_build", "test_cpp_api")
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.strip().endswith("ok")


def test_cpp_statistics():
    """tests/cpp/test_statistics.cpp: the two-constraint greedy policy (exact LP vertices) and the
    progressive widening of the UCT statistic"""
    oracle.build()
    exe = os.path.join(oracle.ORACLE_DIR, "_build", "test_statistics")
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.strip().endswith("ok")


@pytest.mark.gpu
@pytest.mark.skipif(not has_gpu(), reason="needs a CUDA device")
def test_example_scripts_run():
    """the examples use the product path only (no CPU backend): they need the GPU"""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "examples", "plan_merge.py"),
                          "--iterations", "200", "--steps", "2"],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("macro action") == 2
    res = subprocess.run([sys.executable, os.path.join(root, "examples", "risk_constrained_rsbg.py"),
                          "--iterations", "256", "--nheuristic"],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("expected risk") == 2 and "model evaluated" in res.stdout
