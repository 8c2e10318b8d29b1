"""The C++ façade (include/se3gcopter_b200.hpp) mirrors the reference's SE3GCOPTER / Trajectory surface:
compile + link a caller shaped like MavGlobalPlanner::plan on CPU, run it on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "facade_demo")


def _build():
    lib = os.path.join(ROOT, "fast_racing_b200")
    cmd = ["g++", "-std=c++14", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "facade_demo.cpp"),
           "-L", lib, "-lfrx", f"-Wl,-rpath,{lib}", "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_facade_compiles_and_links_as_cxx14():
    _build()  # the reference builds with -std=c++14 (plan_manage/CMakeLists.txt:5)
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_facade_runs_the_reference_call_sequence():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Final cost" in r.stdout and "infeasible corridor rejected: 1" in r.stdout


def test_trajectory_containers_cpu():
    """Piece / Trajectory helpers the planner calls downstream: exact max velocity / acceleration rates
    (root isolation instead of root_finder.hpp) against 2e5-point sampling, checkMax*Rate, getRotation, normalised
    coefficients, junction states.  Header-only, C++14, no CUDA and no libfrx needed."""
    exe = os.path.join(ROOT, "tests", "cpp", "trajectory_check")
    cmd = ["g++", "-std=c++14", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "trajectory_check.cpp"),
           "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == "OK", r.stdout + r.stderr
