"""ctypes binding of the C ABI in include/frx.h (libfrx.so) — plumbing only.

The product is the shared library; this module mirrors its entry points one to one so that
tests, bench.py and Python callers can drive it.  It never computes anything itself and it has
no fallback: if libfrx.so is missing or no CUDA device is present, calls raise.
"""
import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FRX_LIB") or os.path.join(_HERE, "libfrx.so")  # FRX_LIB: instrumented debug variant

FRX_OK = 0
FRX_ERR_INVALID_ARG = -1
FRX_ERR_INFEASIBLE_GEOMETRY = -2
FRX_ERR_CUDA = -3
FRX_ERR_CAPACITY = -4


class FrxParams(C.Structure):
    """frx_params — scalar arguments of SE3GCOPTER::setup (se3gcopter_cpu.hpp:1076-1092)."""

    _fields_ = [
        ("rho", C.c_double),
        ("total_t", C.c_double),
        ("grid_res", C.c_double),
        ("itg_spaces", C.c_int),
        ("c2_diffeo", C.c_int),
        ("horiz_half_len", C.c_double),
        ("vert_half_len", C.c_double),
        ("safe_margin", C.c_double),
        ("v_max", C.c_double),
        ("thr_acc_min", C.c_double),
        ("thr_acc_max", C.c_double),
        ("body_rate_max", C.c_double),
        ("grav_acc", C.c_double),
        ("chi", C.c_double * 4),
    ]


class FrxProblem(C.Structure):
    _fields_ = [
        ("ini", C.c_double * 9),
        ("fin", C.c_double * 9),
        ("n_poly", C.c_int),
        ("face_off", C.POINTER(C.c_int)),
        ("faces", C.POINTER(C.c_double)),
        ("vert_off", C.POINTER(C.c_int)),
        ("verts", C.POINTER(C.c_double)),
        ("params", FrxParams),
    ]


# misc/zhangjiajie_params.yaml of the reference
ZHANGJIAJIE = dict(rho=1000.0, total_t=0.0, grid_res=math.inf, itg_spaces=48, c2_diffeo=1, horiz_half_len=0.5,
                   vert_half_len=0.15, safe_margin=0.08, v_max=14.0, thr_acc_min=5.0, thr_acc_max=12.0,
                   body_rate_max=3.8, grav_acc=9.81, chi=(1.0e7, 1.0e4, 1.0e4, 1.0e4), rel_tol=1.0e-6)
# misc/moderncity_params.yaml: VelMax 21, ThrustAccMax 18, OptRelTol 1e-5
MODERNCITY = dict(ZHANGJIAJIE, v_max=21.0, thr_acc_max=18.0, rel_tol=1.0e-5)


def make_params(d):
    p = FrxParams()
    for k, v in d.items():
        if k == "chi":
            p.chi = (C.c_double * 4)(*v)
        elif k == "rel_tol":
            continue
        else:
            setattr(p, k, v)
    return p


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Problem:
    """Host-side description of one candidate trajectory (keeps the numpy arrays alive)."""

    def __init__(self, ini, fin, face_off, faces, params, vert_off=None, verts=None):
        self.ini = np.ascontiguousarray(ini, dtype=np.float64).reshape(9)
        self.fin = np.ascontiguousarray(fin, dtype=np.float64).reshape(9)
        self.face_off = np.ascontiguousarray(face_off, dtype=np.int32)
        self.faces = np.ascontiguousarray(faces, dtype=np.float64).reshape(-1)
        self.params = dict(params)
        self.vert_off = None if vert_off is None else np.ascontiguousarray(vert_off, dtype=np.int32)
        self.verts = None if verts is None else np.ascontiguousarray(verts, dtype=np.float64).reshape(-1)

    @property
    def n_poly(self):
        return len(self.face_off) - 1

    def with_vpolys(self, vert_off, verts):
        return Problem(self.ini, self.fin, self.face_off, self.faces, self.params, vert_off, verts)

    def as_struct(self):
        s = FrxProblem()
        s.ini = (C.c_double * 9)(*self.ini)
        s.fin = (C.c_double * 9)(*self.fin)
        s.n_poly = self.n_poly
        s.face_off = _ip(self.face_off)
        s.faces = _dp(self.faces)
        if self.vert_off is not None:
            s.vert_off = _ip(self.vert_off)
            s.verts = _dp(self.verts)
        s.params = make_params(self.params)
        return s


_lib = None


def load_library():
    """Loads libfrx.so (raises if it has not been built: there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libfrx.so not built (run `python -c 'import __graft_entry__ as g; g.build()'`): "
                           "there is no CPU fallback for the CUDA path")
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    sig = {
        "frx_batch_create": (ci, [P(FrxProblem), ci, ci, P(vp)]),
        "frx_batch_destroy": (None, [vp]),
        "frx_batch_count": (ci, [vp]),
        "frx_batch_dims": (ci, [vp, ci, P(ci), P(ci), P(ci), P(ci)]),
        "frx_batch_x_offsets": (ci, [vp, P(ci)]),
        "frx_batch_piece_offsets": (ci, [vp, P(ci)]),
        "frx_batch_tables": (ci, [vp, ci, P(ci), P(ci), P(ci), P(ci)]),
        "frx_batch_vpolys": (ci, [vp, ci, P(ci), P(cd)]),
        "frx_batch_initial_guess": (ci, [vp, P(cd)]),
        "frx_batch_eval": (ci, [vp, P(cd), P(cd), P(cd)]),
        "frx_batch_eval_device": (ci, [vp, vp, vp, vp, vp]),
        "frx_batch_eval_launches": (ci, [vp]),
        "frx_batch_stream": (vp, [vp]),
        "frx_batch_host_buffers": (ci, [vp, P(P(cd)), P(P(cd)), P(P(cd))]),
        "frx_batch_set_reference_order": (ci, [vp, ci]),
        "frx_batch_set_solver_slots": (ci, [vp, ci]),
        "frx_eval": (cd, [vp, P(cd), P(cd), ci]),
        "frx_batch_optimize": (ci, [vp, cd, P(cd), P(cd), P(cd), P(cd), P(ci), P(ci), P(ci), P(cd)]),
        "frx_batch_argmin": (ci, [vp, P(ci), P(cd)]),
        "frx_batch_last_ms": (cd, [vp]),
        "frx_strerror": (C.c_char_p, [ci]),
        "frx_version": (C.c_char_p, []),
        "frx_batch_debug_fetch": (ci, [vp, P(cd), P(cd), P(cd), P(cd), P(cd), P(cd), P(cd)]),
        "frx_enumerate_vertices": (ci, [P(cd), ci, P(cd), ci]),
        "frx_host_model": (ci, [P(FrxProblem), P(ci), P(ci), P(ci), P(ci), P(ci), P(ci), P(ci), P(ci), P(cd)]),
        "frx_host_vpolys": (ci, [P(FrxProblem), P(ci), P(cd)]),
        "frx_batch_solver_ticks": (ci, [vp]),
        "frx_batch_solver_update_stats": (ci, [vp, P(cd)]),
        "frx_batch_solver_direction_check": (ci, [vp, P(cd)]),
        "frx_batch_profile_eval": (ci, [vp, vp, vp, vp, vp, ci, P(cd)]),
        "frx_batch_penalty_launch": (ci, [vp, P(C.c_longlong)]),
        "frx_measure_fp64_tflops": (cd, []),
        "frx_corridor_dilate": (ci, [P(cd), ci, P(cd), P(cd), P(cd), P(cd), ci, P(cd)]),
        "frx_corridor_generate": (ci, [P(cd), ci, P(cd), ci, P(cd), cd, vp, vp, P(ci), P(cd), ci, ci, P(ci)]),
        "frx_cloud_create": (ci, [P(cd), ci, ci, P(vp)]),
        "frx_cloud_destroy": (None, [vp]),
        "frx_cloud_dilate_batch": (ci, [vp, P(cd), ci, P(cd), P(ci), P(cd), P(cd)]),
        "frx_cloud_generate_batch": (ci, [vp, P(cd), P(ci), ci, P(cd), cd, vp, vp, P(ci), P(ci), P(cd), ci, ci]),
        "frx_device_numa_node": (ci, [ci]),
        "frx_bind_thread_to_device": (ci, [ci]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    "frx_batch_create", "frx_batch_destroy", "frx_batch_count", "frx_batch_dims", "frx_batch_x_offsets",
    "frx_batch_piece_offsets", "frx_batch_tables", "frx_batch_vpolys", "frx_batch_initial_guess", "frx_batch_eval",
    "frx_batch_eval_device", "frx_batch_eval_launches", "frx_batch_stream", "frx_batch_host_buffers", "frx_batch_set_reference_order", "frx_eval", "frx_batch_optimize", "frx_batch_argmin",
    "frx_batch_last_ms", "frx_strerror", "frx_version", "frx_device_numa_node", "frx_bind_thread_to_device",
    "frx_batch_argmin_allgather", "frx_batch_set_solver_slots", "frx_corridor_dilate", "frx_corridor_generate", "frx_cloud_create", "frx_cloud_destroy",
    "frx_cloud_dilate_batch", "frx_cloud_generate_batch",
]


class FrxError(RuntimeError):
    def __init__(self, status, what):
        lib = load_library()
        super().__init__(f"{what}: {lib.frx_strerror(status).decode()} ({status})")
        self.status = status


def _check(status, what):
    if status != FRX_OK:
        raise FrxError(status, what)


def host_model(problem):
    """Index tables and starting point of one problem, host only (no GPU needed)."""
    lib = load_library()
    s = problem.as_struct()
    n, M, dt, dpp = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    _check(lib.frx_host_model(C.byref(s), C.byref(n), C.byref(M), C.byref(dt), C.byref(dpp), None, None, None, None, None),
           "frx_host_model")
    intervals = np.zeros(problem.n_poly, np.int32)
    idx_hs = np.zeros(M.value, np.int32)
    idx_vs = np.zeros(max(M.value - 1, 1), np.int32)
    cons = np.zeros(M.value, np.int32)
    x0 = np.zeros(n.value)
    _check(lib.frx_host_model(C.byref(s), None, None, None, None, _ip(intervals), _ip(idx_hs), _ip(idx_vs), _ip(cons), _dp(x0)),
           "frx_host_model")
    return dict(n=n.value, pieces=M.value, dim_t=dt.value, dim_p=dpp.value, intervals=intervals, idx_hs=idx_hs,
                idx_vs=idx_vs[: M.value - 1], cons=cons, x0=x0)


def host_vpolys(problem):
    lib = load_library()
    s = problem.as_struct()
    off = np.zeros(2 * problem.n_poly, np.int32)
    _check(lib.frx_host_vpolys(C.byref(s), _ip(off), None), "frx_host_vpolys")
    verts = np.zeros(3 * int(off[-1]))
    _check(lib.frx_host_vpolys(C.byref(s), _ip(off), _dp(verts)), "frx_host_vpolys")
    return off, verts


def enumerate_vertices(faces):
    """Vertices (V x 3) of one H-polytope given as F x 6 (normal, point) rows; host only."""
    lib = load_library()
    f = np.ascontiguousarray(faces, dtype=np.float64).reshape(-1)
    out = np.zeros(3 * 512)
    v = lib.frx_enumerate_vertices(_dp(f), len(f) // 6, _dp(out), 512)
    if v < 0:
        raise FrxError(v, "frx_enumerate_vertices")
    return out[: 3 * v].reshape(v, 3).copy()


CORRIDOR_MAX_FACES = 96
IS_BLOCKED_T = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double))


def corridor_dilate(obs, p1, p2, bbox):
    """LineSegment3D::dilate for one segment on the host: (faces F x 6 as (n, p), ellipsoid C 3x3, d)."""
    lib = load_library()
    obs = np.ascontiguousarray(obs, dtype=np.float64).reshape(-1)
    p1, p2, bbox = (np.ascontiguousarray(v, dtype=np.float64) for v in (p1, p2, bbox))
    faces = np.zeros(6 * 512)
    ell = np.zeros(12)
    F = lib.frx_corridor_dilate(_dp(obs), len(obs) // 3, _dp(p1), _dp(p2), _dp(bbox), _dp(faces), 512, _dp(ell))
    if F < 0:
        raise FrxError(F, "frx_corridor_dilate")
    return faces[: 6 * F].reshape(F, 6).copy(), ell[:9].reshape(3, 3).copy(), ell[9:].copy()


def corridor_generate(obs, path, bbox, map_height, blocked=None):
    """The walk of MavGlobalPlanner::plan along `path` (host): returns (face_off, faces F x 6); blocked(a, b) -> bool or None."""
    lib = load_library()
    obs = np.ascontiguousarray(obs, dtype=np.float64).reshape(-1)
    path = np.ascontiguousarray(path, dtype=np.float64).reshape(-1)
    bbox = np.ascontiguousarray(bbox, dtype=np.float64)
    cap_poly, cap_faces = len(path) // 3 + 1, (len(path) // 3 + 1) * CORRIDOR_MAX_FACES
    off = np.zeros(cap_poly + 1, np.int32)
    faces = np.zeros(6 * cap_faces)
    n = C.c_int()
    cb = IS_BLOCKED_T(lambda ctx, a, b: int(bool(blocked(np.array(a[:3]), np.array(b[:3]))))) if blocked else None
    st = lib.frx_corridor_generate(_dp(obs), len(obs) // 3, _dp(path), len(path) // 3, _dp(bbox), map_height,
                                   C.cast(cb, C.c_void_p) if cb else None, None, _ip(off), _dp(faces), cap_poly, cap_faces, C.byref(n))
    _check(st, "frx_corridor_generate")
    off = off[: n.value + 1].copy()
    return off, faces[: 6 * off[-1]].reshape(-1, 6).copy()


class Cloud:
    """An obstacle cloud resident on one GPU (frx_cloud): the segment step of the corridor generation, batched."""

    def __init__(self, obs, device=-1):
        self.lib = load_library()
        obs = np.ascontiguousarray(obs, dtype=np.float64).reshape(-1)
        h = C.c_void_p()
        _check(self.lib.frx_cloud_create(_dp(obs), len(obs) // 3, device, C.byref(h)), "frx_cloud_create")
        self.h = h

    def dilate_batch(self, segments, bbox):
        seg = np.ascontiguousarray(segments, dtype=np.float64).reshape(-1, 6)
        bbox = np.ascontiguousarray(bbox, dtype=np.float64)
        n = len(seg)
        cnt = np.zeros(n, np.int32)
        faces = np.zeros(n * CORRIDOR_MAX_FACES * 6)
        ell = np.zeros(12 * n)
        _check(self.lib.frx_cloud_dilate_batch(self.h, _dp(seg.reshape(-1)), n, _dp(bbox), _ip(cnt), _dp(faces), _dp(ell)), "frx_cloud_dilate_batch")
        faces = faces.reshape(n, CORRIDOR_MAX_FACES, 6)
        return [faces[i, : max(cnt[i], 0)].copy() for i in range(n)], cnt, ell.reshape(n, 12)

    def generate_batch(self, paths, bbox, map_height):
        """Corridors of many candidate paths at once (lock-step walks, one batched device step per walk step):
        returns a list of (face_off, faces F x 6) per path."""
        bbox = np.ascontiguousarray(bbox, dtype=np.float64)
        poff = np.zeros(len(paths) + 1, np.int32)
        for i, p in enumerate(paths):
            poff[i + 1] = poff[i] + len(p)
        flat = np.ascontiguousarray(np.vstack(paths), dtype=np.float64).reshape(-1)
        cap_poly, cap_faces = int(poff[-1]), int(poff[-1]) * CORRIDOR_MAX_FACES
        poly_off = np.zeros(len(paths) + 1, np.int32)
        face_off = np.zeros(cap_poly + 1, np.int32)
        faces = np.zeros(6 * cap_faces)
        _check(self.lib.frx_cloud_generate_batch(self.h, _dp(flat), _ip(poff), len(paths), _dp(bbox), map_height, None, None, _ip(poly_off),
                                                 _ip(face_off), _dp(faces), cap_poly, cap_faces), "frx_cloud_generate_batch")
        out = []
        for i in range(len(paths)):
            fo = face_off[poly_off[i]:poly_off[i + 1] + 1]
            out.append((fo - fo[0], faces[6 * fo[0]:6 * fo[-1]].reshape(-1, 6).copy()))
        return out

    def close(self):
        if getattr(self, "h", None):
            self.lib.frx_cloud_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    """A batch of candidate trajectories resident on one GPU (frx_batch)."""

    def __init__(self, problems, device=-1):
        self.lib = load_library()
        self.problems = list(problems)
        arr = (FrxProblem * len(self.problems))(*[p.as_struct() for p in self.problems])
        h = C.c_void_p()
        _check(self.lib.frx_batch_create(arr, len(self.problems), device, C.byref(h)), "frx_batch_create")
        self.h = h
        B = len(self.problems)
        self.count = B
        self.x_off = np.zeros(B + 1, np.int32)
        self.piece_off = np.zeros(B + 1, np.int32)
        _check(self.lib.frx_batch_x_offsets(h, _ip(self.x_off)), "x_offsets")
        _check(self.lib.frx_batch_piece_offsets(h, _ip(self.piece_off)), "piece_offsets")
        self.total_x = int(self.x_off[-1])
        self.total_pieces = int(self.piece_off[-1])

    def close(self):
        if getattr(self, "h", None):
            self.lib.frx_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dims(self, traj=0):
        n, M, dt, dpp = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _check(self.lib.frx_batch_dims(self.h, traj, C.byref(n), C.byref(M), C.byref(dt), C.byref(dpp)), "dims")
        return dict(n=n.value, pieces=M.value, dim_t=dt.value, dim_p=dpp.value)

    def tables(self, traj=0):
        d = self.dims(traj)
        npoly = self.problems[traj].n_poly
        intervals = np.zeros(npoly, np.int32)
        idx_hs = np.zeros(d["pieces"], np.int32)
        idx_vs = np.zeros(max(d["pieces"] - 1, 1), np.int32)
        cons = np.zeros(d["pieces"], np.int32)
        _check(self.lib.frx_batch_tables(self.h, traj, _ip(intervals), _ip(idx_hs), _ip(idx_vs), _ip(cons)), "tables")
        return dict(intervals=intervals, idx_hs=idx_hs, idx_vs=idx_vs[: d["pieces"] - 1], cons=cons)

    def vpolys(self, traj=0):
        npoly = self.problems[traj].n_poly
        off = np.zeros(2 * npoly, np.int32)
        _check(self.lib.frx_batch_vpolys(self.h, traj, _ip(off), None), "vpolys")
        verts = np.zeros(3 * int(off[-1]))
        _check(self.lib.frx_batch_vpolys(self.h, traj, _ip(off), _dp(verts)), "vpolys")
        return off, verts

    def initial_guess(self):
        x = np.zeros(self.total_x)
        _check(self.lib.frx_batch_initial_guess(self.h, _dp(x)), "initial_guess")
        return x

    def eval(self, x):
        """Host-buffer evaluation (H2D + kernels + D2H): returns (cost[B], grad[total_x])."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == self.total_x
        cost = np.zeros(self.count)
        grad = np.zeros(self.total_x)
        _check(self.lib.frx_batch_eval(self.h, _dp(x), _dp(cost), _dp(grad)), "frx_batch_eval")
        return cost, grad

    def set_reference_order(self, on=True):
        """Reference-order mode: bit-identical to the reference CPU code, slow (validation / reproducibility)."""
        _check(self.lib.frx_batch_set_reference_order(self.h, 1 if on else 0), "set_reference_order")

    def set_solver_slots(self, slots):
        """Queue-fed solver slots (0 = one per trajectory): see frx_batch_set_solver_slots."""
        _check(self.lib.frx_batch_set_solver_slots(self.h, int(slots)), "set_solver_slots")

    def host_buffers(self):
        """numpy views of the batch's pinned host buffers (x, cost, grad): eval_pinned() transfers straight from/to them."""
        px, pc, pg = C.POINTER(C.c_double)(), C.POINTER(C.c_double)(), C.POINTER(C.c_double)()
        _check(self.lib.frx_batch_host_buffers(self.h, C.byref(px), C.byref(pc), C.byref(pg)), "host_buffers")
        self._pins = (px, pc, pg)
        return (np.ctypeslib.as_array(px, shape=(self.total_x,)), np.ctypeslib.as_array(pc, shape=(self.count,)),
                np.ctypeslib.as_array(pg, shape=(self.total_x,)))

    def eval_pinned(self):
        """frx_batch_eval on the batch's own pinned buffers (fill host_buffers()[0] first): no staging copies."""
        px, pc, pg = self._pins
        _check(self.lib.frx_batch_eval(self.h, px, pc, pg), "frx_batch_eval")

    def eval_device(self, x_ptr, cost_ptr, grad_ptr, stream_ptr=0):
        """Device-pointer evaluation on a CUDA stream (integers: data_ptr() / cuda_stream)."""
        _check(self.lib.frx_batch_eval_device(self.h, C.c_void_p(x_ptr), C.c_void_p(cost_ptr), C.c_void_p(grad_ptr),
                                              C.c_void_p(stream_ptr)), "frx_batch_eval_device")

    def stream(self):
        """cudaStream_t (as int) of the batch's own stream."""
        return int(self.lib.frx_batch_stream(self.h) or 0)

    def profile_eval(self, x_ptr, cost_ptr, grad_ptr, stream_ptr, reps):
        """Average device ms of (minco_forward, penalty, minco_backward) over reps evaluations."""
        ms = np.zeros(3)
        _check(self.lib.frx_batch_profile_eval(self.h, C.c_void_p(x_ptr), C.c_void_p(cost_ptr), C.c_void_p(grad_ptr),
                                               C.c_void_p(stream_ptr), reps, _dp(ms)), "frx_batch_profile_eval")
        return ms

    def penalty_launch(self):
        out = (C.c_longlong * 6)()
        _check(self.lib.frx_batch_penalty_launch(self.h, out), "penalty_launch")
        return dict(zip(("samples_per_thread", "threads_per_piece", "pieces_per_cta", "cta_threads", "grid", "smem_bytes"), list(out)))

    def eval_launches(self):
        return self.lib.frx_batch_eval_launches(self.h)

    def debug_fetch(self):
        TP = self.total_pieces
        out = {k: np.zeros(s) for k, s in dict(T=TP, q=3 * TP, c=18 * TP, lam=18 * TP, pen=TP, energy=TP, gdT=TP).items()}
        _check(self.lib.frx_batch_debug_fetch(self.h, _dp(out["T"]), _dp(out["q"]), _dp(out["c"]), _dp(out["lam"]),
                                              _dp(out["pen"]), _dp(out["energy"]), _dp(out["gdT"])), "debug_fetch")
        return out

    def optimize(self, rel_tol):
        B, TP = self.count, self.total_pieces
        r = dict(durations=np.zeros(TP), coeffs=np.zeros(18 * TP), jerk_cost=np.zeros(B), final_cost=np.zeros(B),
                 status=np.zeros(B, np.int32), iterations=np.zeros(B, np.int32), evaluations=np.zeros(B, np.int32),
                 x=np.zeros(self.total_x))
        _check(self.lib.frx_batch_optimize(self.h, rel_tol, _dp(r["durations"]), _dp(r["coeffs"]), _dp(r["jerk_cost"]),
                                           _dp(r["final_cost"]), _ip(r["status"]), _ip(r["iterations"]),
                                           _ip(r["evaluations"]), _dp(r["x"])), "frx_batch_optimize")
        r["device_ms"] = self.last_ms()
        r["ticks"] = self.lib.frx_batch_solver_ticks(self.h)
        st = np.zeros(3)
        self.lib.frx_batch_solver_update_stats(self.h, _dp(st))
        chk = C.c_double(-1.0)
        self.lib.frx_batch_solver_direction_check(self.h, C.byref(chk))
        r["direction_check"] = chk.value
        r["update_kernel"] = {"sampled_launches": int(st[2]), "device_ms": float(st[0]), "history_bytes": float(st[1])}
        return r

    def argmin(self):
        i, v = C.c_int(), C.c_double()
        _check(self.lib.frx_batch_argmin(self.h, C.byref(i), C.byref(v)), "frx_batch_argmin")
        return i.value, v.value

    def last_ms(self):
        return self.lib.frx_batch_last_ms(self.h)

    def eval_callback(self):
        """(callable, instance) usable as an lbfgs_evaluate_t: frx_eval on a batch of ONE trajectory."""
        assert self.count == 1
        return self.lib.frx_eval, self.h
