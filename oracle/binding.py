"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/liboracle.so (and oracle/_ref).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; nothing under fast_racing_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF_LBFGS = os.path.join(HERE, "_ref", "libref_lbfgs.so")
REF_GCOPTER = os.path.join(HERE, "_ref", "libfrx_ref.so")
REF_GCOPTER_FULL = os.path.join(HERE, "_ref", "libfrx_ref_full.so")  # + verbatim geoutils.hpp / sdlp.hpp / quickhull.hpp
REF_GEO = os.path.join(HERE, "_ref", "libref_geo.so")

EVAL_T = C.CFUNCTYPE(C.c_double, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int)


def build():
    subprocess.run(["make", "-C", HERE], check=True, capture_output=True)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        P, vp, ci, cd = C.POINTER, C.c_void_p, C.c_int, C.c_double
        L.fro_create.restype = vp
        L.fro_create.argtypes = [vp]
        L.fro_destroy.argtypes = [vp]
        L.fro_dims.argtypes = [vp, P(ci), P(ci), P(ci), P(ci)]
        L.fro_tables.argtypes = [vp, P(ci), P(ci), P(ci), P(ci)]
        L.fro_initial_guess.argtypes = [vp, P(cd)]
        L.fro_eval.restype = cd
        L.fro_eval.argtypes = [vp, P(cd), P(cd), ci]
        L.fro_forward.argtypes = [vp, P(cd), P(cd), P(cd), P(cd)]
        L.fro_penalty_only.argtypes = [vp, P(cd), P(cd), P(cd)]
        L.fro_optimize.restype = cd
        L.fro_optimize.argtypes = [vp, cd, P(cd), P(cd), P(cd), P(cd), P(ci), P(ci), P(ci)]
        L.fro_optimize_from.restype = cd
        L.fro_optimize_from.argtypes = [vp, P(cd), cd]
        L.fro_time_evals.restype = cd
        L.fro_time_evals.argtypes = [vp, P(cd), ci]
        L.fro_time_evals_mt.restype = cd
        L.fro_time_evals_mt.argtypes = [P(vp), P(P(cd)), ci, ci, ci]
        L.fro_time_solves_mt.restype = cd
        L.fro_time_solves_mt.argtypes = [P(vp), ci, cd, ci, P(cd), P(ci), P(ci), P(ci)]
        L.fro_brute_force_vertices.argtypes = [P(cd), ci, P(cd), ci]
        L.fro_vpolys.argtypes = [vp, P(ci), P(cd)]
        L.fro_set_max_iterations.argtypes = [vp, ci]
        L.fro_banded_solve.argtypes = [P(cd), ci, P(cd), ci]
        L.fro_lbfgs_testfn.argtypes = [ci, ci, P(cd), P(cd), ci, ci, cd, cd, ci, P(ci), P(ci)]
        L.fro_testfn_eval.restype = cd
        L.fro_testfn_eval.argtypes = [ci, ci, P(cd), P(cd)]
        _lib = L
    return _lib


class Oracle:
    """CPU oracle for one problem.  `problem` must carry V-polytopes (Problem.with_vpolys)."""

    def __init__(self, problem):
        self.L = lib()
        assert problem.vert_off is not None, "the oracle takes the V-polytopes as input"
        self.problem = problem
        self._s = problem.as_struct()
        self.h = self.L.fro_create(C.cast(C.pointer(self._s), C.c_void_p))
        if not self.h:
            raise RuntimeError("oracle setup failed")
        n, M, dt, dpp = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self.L.fro_dims(self.h, C.byref(n), C.byref(M), C.byref(dt), C.byref(dpp))
        self.n, self.pieces, self.dim_t, self.dim_p = n.value, M.value, dt.value, dpp.value

    def close(self):
        if self.h:
            self.L.fro_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tables(self):
        intervals = np.zeros(self.problem.n_poly, np.int32)
        idx_hs = np.zeros(self.pieces, np.int32)
        idx_vs = np.zeros(max(self.pieces - 1, 1), np.int32)
        cons = np.zeros(self.pieces, np.int32)
        self.L.fro_tables(self.h, _ip(intervals), _ip(idx_hs), _ip(idx_vs), _ip(cons))
        return dict(intervals=intervals, idx_hs=idx_hs, idx_vs=idx_vs[: self.pieces - 1], cons=cons)

    def initial_guess(self):
        x = np.zeros(self.n)
        self.L.fro_initial_guess(self.h, _dp(x))
        return x

    def eval(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        g = np.zeros(self.n)
        f = self.L.fro_eval(self.h, _dp(x), _dp(g), self.n)
        return f, g

    def forward(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        T = np.zeros(self.pieces)
        q = np.zeros(3 * max(self.pieces - 1, 1))
        c = np.zeros(18 * self.pieces)
        self.L.fro_forward(self.h, _dp(x), _dp(T), _dp(q), _dp(c))
        return T, q[: 3 * (self.pieces - 1)], c

    def penalty_only(self):
        cost = C.c_double()
        gdT = np.zeros(self.pieces)
        gdC = np.zeros(18 * self.pieces)
        self.L.fro_penalty_only(self.h, C.byref(cost), _dp(gdT), _dp(gdC))
        return cost.value, gdT, gdC

    def optimize(self, rel_tol):
        dur = np.zeros(self.pieces)
        co = np.zeros(18 * self.pieces)
        xf = np.zeros(self.n)
        fc, st, it, ev = C.c_double(), C.c_int(), C.c_int(), C.c_int()
        j = self.L.fro_optimize(self.h, rel_tol, _dp(dur), _dp(co), _dp(xf), C.byref(fc), C.byref(st), C.byref(it), C.byref(ev))
        return dict(jerk_cost=j, durations=dur, coeffs=co, x=xf, final_cost=fc.value, status=st.value,
                    iterations=it.value, evaluations=ev.value)

    def set_max_iterations(self, n):
        """Safety net of timing runs (0 = no limit, as the reference): a NaN objective otherwise never stops."""
        self.L.fro_set_max_iterations(self.h, int(n))

    def time_evals(self, x, reps):
        x = np.ascontiguousarray(x, dtype=np.float64)
        return self.L.fro_time_evals(self.h, _dp(x), reps)


def time_evals_mt(oracles, xs, reps, threads):
    L = lib()
    n = len(oracles)
    hs = (C.c_void_p * n)(*[o.h for o in oracles])
    xs = [np.ascontiguousarray(x, dtype=np.float64) for x in xs]
    xp = (C.POINTER(C.c_double) * n)(*[_dp(x) for x in xs])
    return L.fro_time_evals_mt(hs, xp, n, reps, threads)


def time_solves_mt(oracles, rel_tol, threads):
    L = lib()
    n = len(oracles)
    hs = (C.c_void_p * n)(*[o.h for o in oracles])
    fc = np.zeros(n)
    st = np.zeros(n, np.int32)
    it = np.zeros(n, np.int32)
    ev = np.zeros(n, np.int32)
    secs = L.fro_time_solves_mt(hs, n, rel_tol, threads, _dp(fc), _ip(st), _ip(it), _ip(ev))
    return secs, dict(final_cost=fc, status=st, iterations=it, evaluations=ev)


def brute_force_vertices(faces):
    L = lib()
    f = np.ascontiguousarray(faces, dtype=np.float64).reshape(-1)
    out = np.zeros(3 * 1024)
    v = L.fro_brute_force_vertices(_dp(f), len(f) // 6, _dp(out), 1024)
    return out[: 3 * v].reshape(v, 3).copy()


def vpolys(problem):
    """V-polytopes of a corridor by the oracle's own brute-force enumeration (no product code involved)."""
    L = lib()
    s = problem.as_struct()
    off = np.zeros(2 * problem.n_poly, np.int32)
    if L.fro_vpolys(C.cast(C.pointer(s), C.c_void_p), _ip(off), None) != 0:
        raise RuntimeError("oracle: empty polytope or intersection")
    verts = np.zeros(3 * int(off[-1]))
    L.fro_vpolys(C.cast(C.pointer(s), C.c_void_p), _ip(off), _dp(verts))
    return off, verts


def ref_lbfgs():
    """Verbatim reference lbfgs.hpp behind a C wrapper (None if oracle/_ref was never built)."""
    if not os.path.exists(REF_LBFGS):
        return None
    R = C.CDLL(REF_LBFGS)
    R.ref_lbfgs_optimize.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), EVAL_T, C.c_void_p, C.c_int,
                                     C.c_int, C.c_double, C.c_double, C.c_double, C.c_int]
    return R


class RefGcopter:
    """The reference's own SE3GCOPTER (verbatim se3gcopter_cpu.hpp + trajectory.hpp + lbfgs.hpp compiled over the
    Eigen-API shim, oracle/_ref/libfrx_ref.so).  `problem` must carry V-polytopes.  Available only where oracle/_ref was
    built (a box that has /root/reference) or travelled with the snapshot."""

    _lib = None
    PATH = REF_GCOPTER
    NEEDS_VPOLYS = True

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)

    @classmethod
    def lib(cls):
        if cls._lib is None:
            R = C.CDLL(cls.PATH)
            P, vp, ci, cd = C.POINTER, C.c_void_p, C.c_int, C.c_double
            R.ref_create.restype = vp
            R.ref_create.argtypes = [vp]
            R.ref_destroy.argtypes = [vp]
            R.ref_dims.argtypes = [vp, P(ci), P(ci), P(ci), P(ci)]
            R.ref_tables.argtypes = [vp, P(ci), P(ci), P(ci), P(ci)]
            R.ref_eval.restype = cd
            R.ref_eval.argtypes = [vp, P(cd), P(cd), ci]
            R.ref_initial_guess.argtypes = [vp, P(cd)]
            R.ref_forward.argtypes = [vp, P(cd), P(cd), P(cd)]
            R.ref_optimize.restype = cd
            R.ref_optimize.argtypes = [vp, cd, P(cd), P(cd)]
            R.ref_vpolys.argtypes = [vp, P(ci), P(cd)]
            cls._lib = R
        return cls._lib

    def __init__(self, problem):
        self.R = self.lib()
        assert problem.vert_off is not None or not self.NEEDS_VPOLYS
        self.problem = problem
        self._s = problem.as_struct()
        self.h = self.R.ref_create(C.cast(C.pointer(self._s), C.c_void_p))
        if not self.h:
            raise RuntimeError("reference SE3GCOPTER::setup returned false")
        n, M, dt, dpp = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self.R.ref_dims(self.h, C.byref(n), C.byref(M), C.byref(dt), C.byref(dpp))
        self.n, self.pieces, self.dim_t, self.dim_p = n.value, M.value, dt.value, dpp.value

    def close(self):
        if self.h:
            self.R.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tables(self):
        intervals = np.zeros(self.problem.n_poly, np.int32)
        idx_hs = np.zeros(self.pieces, np.int32)
        idx_vs = np.zeros(max(self.pieces - 1, 1), np.int32)
        cons = np.zeros(self.pieces, np.int32)
        self.R.ref_tables(self.h, _ip(intervals), _ip(idx_hs), _ip(idx_vs), _ip(cons))
        return dict(intervals=intervals, idx_hs=idx_hs, idx_vs=idx_vs[: self.pieces - 1], cons=cons)

    def initial_guess(self):
        x = np.zeros(self.n)
        self.R.ref_initial_guess(self.h, _dp(x))
        return x

    def eval(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        g = np.zeros(self.n)
        f = self.R.ref_eval(self.h, _dp(x), _dp(g), self.n)
        return f, g

    def forward(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        T = np.zeros(self.pieces)
        c = np.zeros(18 * self.pieces)
        self.R.ref_forward(self.h, _dp(x), _dp(T), _dp(c))
        return T, c

    def optimize(self, rel_tol):
        dur = np.zeros(self.pieces)
        co = np.zeros(18 * self.pieces)
        j = self.R.ref_optimize(self.h, rel_tol, _dp(dur), _dp(co))
        return dict(jerk_cost=j, durations=dur, coeffs=co)

    def vpolys(self):
        """The V-polytopes the reference's setup holds (cfgVs), origin + edge columns, in the reference's order."""
        off = np.zeros(2 * self.problem.n_poly, np.int32)
        self.R.ref_vpolys(self.h, _ip(off), None)
        verts = np.zeros(3 * int(off[-1]))
        self.R.ref_vpolys(self.h, _ip(off), _dp(verts))
        return off, verts


class RefGcopterFull(RefGcopter):
    """The reference's SE3GCOPTER with NOTHING supplied: verbatim geoutils.hpp + sdlp.hpp + quickhull.hpp too, so setup() runs
    the reference's own interior-point LP and dual quickhull (oracle/_ref/libfrx_ref_full.so).  One fresh process per use is
    advisable when bit-reproducibility matters: sdlp keeps a function-static RNG (sdlp.hpp:693)."""

    _lib = None
    PATH = REF_GCOPTER_FULL
    NEEDS_VPOLYS = False


def ref_enumerate_vertices(faces):
    """Vertices (V x 3, absolute, the reference's order) of one H-polytope by the verbatim geoutils::enumerateVs; None if empty."""
    if not os.path.exists(REF_GEO):
        return None
    G = C.CDLL(REF_GEO)
    f = np.ascontiguousarray(faces, dtype=np.float64).reshape(-1)
    out = np.zeros(3 * 1024)
    G.ref_enumerate_vertices.argtypes = [C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_int]
    v = G.ref_enumerate_vertices(_dp(f), len(f) // 6, _dp(out), 1024)
    return None if v < 0 else out[: 3 * v].reshape(v, 3).copy()
