"""Launch-geometry rules of the site-parallel engine, checked on the host (no GPU): tests/cpp/geometry_check.cu
sweeps models x visits x covariates x dtype x batch size through csrc/engine.cuh:plan_geometry."""

import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_plan_geometry_invariants(tmp_path):
    exe = tmp_path / "geometry_check"
    src = os.path.join(ROOT, "tests", "cpp", "geometry_check.cu")
    cc = subprocess.run(["nvcc", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(exe), src],
                        capture_output=True, text=True, timeout=600)
    assert cc.returncode == 0, cc.stderr[-2000:]
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout[-4000:]
    assert "failed=0" in run.stdout
