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


def _build_eigen_overloads():
    lib = os.path.join(ROOT, "fast_racing_b200")
    exe = os.path.join(ROOT, "tests", "cpp", "facade_eigen_overloads")
    cmd = ["g++", "-std=c++14", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle", "eigen_shim"),
           os.path.join(ROOT, "tests", "cpp", "facade_eigen_overloads.cpp"), "-L", lib, "-lfrx", f"-Wl,-rpath,{lib}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_facade_eigen_overloads_compile():
    """The reference's exact call (Eigen::MatrixXd hPolys, Matrix3d states, Vector4d weights, MinCoPlan_CPU.cpp:114-124)
    compiles against the façade's Eigen overloads; oracle/eigen_shim stands in for Eigen (not in this image)."""
    assert os.path.exists(_build_eigen_overloads())


@pytest.mark.gpu
def test_facade_eigen_overloads_run():
    r = subprocess.run([_build_eigen_overloads()], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_c_abi_cross_rank_argmin_nccl():
    """frx_batch_argmin_allgather from plain C++ (tests/cpp/argmin_nccl.cpp): one thread per visible GPU, ncclCommInitAll,
    every rank must agree on the (cost, global index) of the cheapest candidate.  With one GPU it is a 1-rank communicator."""
    lib = os.path.join(ROOT, "fast_racing_b200")
    exe = os.path.join(ROOT, "tests", "cpp", "argmin_nccl")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-std=c++14", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           os.path.join(ROOT, "tests", "cpp", "argmin_nccl.cpp"), "-L", lib, "-lfrx", "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lnccl",
           "-pthread", f"-Wl,-rpath,{lib}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "all ranks agree" in r.stdout, r.stdout + r.stderr


def test_facade_corridor_generation_cpu():
    """the corridor half of MavGlobalPlanner::plan written against frx::EllipsoidDecomp3D / Polyhedron3D with the
    reference's own call sequence (MinCoPlan_CPU.cpp:37-105): polytopes cover the path ends, no obstacle point is strictly inside."""
    lib = os.path.join(ROOT, "fast_racing_b200")
    exe = os.path.join(ROOT, "tests", "cpp", "facade_corridor")
    cmd = ["g++", "-std=c++14", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "facade_corridor.cpp"),
           "-L", lib, "-lfrx", f"-Wl,-rpath,{lib}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


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


def test_trajectory_containers_match_the_reference_trajectory_hpp():
    """frx::Piece / frx::Trajectory / traj2fields against the reference's OWN trajectory.hpp + root_finder.hpp
    (compiled verbatim over the Eigen-API shim into oracle/_ref/libfrx_ref.so) on the reference's optimised trajectories of
    the golden fixtures: positions / derivatives, getRotation, getTiltRate, total duration, junction states, the
    duration-normalised coefficients that traj2msg ships (se3_planner.cpp:39-53), and the max-rate report of
    MinCoPlan_CPU.cpp:131-132 (the façade isolates roots itself instead of using root_finder.hpp)."""
    import ctypes as C

    import numpy as np

    ref_so = os.path.join(ROOT, "oracle", "_ref", "libfrx_ref.so")
    if not os.path.exists(ref_so):
        pytest.skip("oracle/_ref not built on this box")
    so = os.path.join(ROOT, "tests", "cpp", "facade_traj_capi.so")
    r = subprocess.run(["g++", "-std=c++14", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "cpp", "facade_traj_capi.cpp"), "-o", so], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    F, R = C.CDLL(so), C.CDLL(ref_so)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    from helpers import golden_names, load_golden

    for name in golden_names():
        _, z = load_golden(name)
        dur = np.ascontiguousarray(z["ref_opt_durations"], dtype=np.float64)
        co = np.ascontiguousarray(z["ref_opt_coeffs"], dtype=np.float64)
        M = len(dur)
        total = float(dur.sum())
        ts = np.ascontiguousarray(np.r_[0.0, np.linspace(0.013, 0.987, 41) * total, total])
        out = {}
        for tag, lib, fn in (("ref", R, "ref_traj_query"), ("fac", F, "facade_traj_query")):
            summary, per_t, norm, wire = np.zeros(3), np.zeros(24 * len(ts)), np.zeros(18 * M), np.zeros(19 * M)
            args = [dp(dur), dp(co), C.c_int(M), dp(ts), C.c_int(len(ts)), C.c_double(0.4), C.c_double(9.81), dp(summary), dp(per_t), dp(norm)]
            if tag == "fac":
                args.append(dp(wire))
            getattr(lib, fn)(*args)
            out[tag] = (summary, per_t.reshape(-1, 24), norm.reshape(M, 6, 3), wire.reshape(M, 19))
        (rs, rt, rn, _), (fs, ft, fn_, fw) = out["ref"], out["fac"]
        assert abs(fs[0] - rs[0]) <= 1e-15 * rs[0]
        # max rates: the reference isolates roots to FLT_EPSILON / duration (trajectory.hpp:207), the façade to machine precision
        assert abs(fs[1] - rs[1]) <= 1e-9 * rs[1] and abs(fs[2] - rs[2]) <= 1e-9 * rs[2], (name, fs, rs)
        scale = np.abs(rt).max(axis=0) + 1e-300
        assert (np.abs(ft - rt).max(axis=0) <= 1e-12 * scale).all(), (name, (np.abs(ft - rt).max(axis=0) / scale).max())
        assert np.array_equal(fn_, rn)  # normalizePosCoeffMat: same products in the same order -> bit-identical
        # wire format = normalised coefficients, highest order first, per axis, + the duration (traj2msg)
        for i in range(M):
            assert np.array_equal(fw[i, 0:6], rn[i, :, 0]) and np.array_equal(fw[i, 6:12], rn[i, :, 1]) and np.array_equal(fw[i, 12:18], rn[i, :, 2])
            assert fw[i, 18] == dur[i]
        for lim_scale in (0.9, 1.1):
            res = {}
            for tag, lib, fn in (("ref", R, "ref_traj_checks"), ("fac", F, "facade_traj_checks")):
                vo, ao, pva = C.c_int(), C.c_int(), np.zeros(9)
                getattr(lib, fn)(dp(dur), dp(co), C.c_int(M), C.c_double(rs[1] * lim_scale), C.c_double(rs[2] * lim_scale), C.c_int(M // 2),
                                 C.byref(vo), C.byref(ao), dp(pva))
                res[tag] = (vo.value, ao.value, pva)
            assert res["ref"][:2] == res["fac"][:2] == ((0, 0) if lim_scale < 1 else (1, 1))
            assert np.array_equal(res["ref"][2], res["fac"][2])
