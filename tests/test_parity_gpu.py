"""GPU tier: the CUDA path (through the C ABI of libfrx.so) against the CPU oracle, the committed golden
vectors, and size-independent properties at BASELINE.json's full sizes.

Tolerances (fp64; stated in DESIGN.md "Numerics"):
  integer tables            bit-exact
  single evaluation         |cost - cost_ref| <= 1e-11 |cost_ref| ; max|grad - grad_ref| <= 1e-10 max|grad_ref|
                            (the kernels use FMA contraction and fixed-order parallel reductions; the oracle sums
                             sequentially without FMA — observed agreement is ~1e-14..1e-12)
  converged solve           cost within 1e-6 relative of the oracle when both run to a tight tolerance
"""
import ctypes as C
import math

import numpy as np
import pytest

from fast_racing_b200 import api, synth
from helpers import golden_names, load_golden, rel_err

pytestmark = pytest.mark.gpu

COST_RTOL = 1e-11
GRAD_RTOL = 1e-10


def _oracles(oracle_mod, batch, problems, idx=None):
    out = []
    for i in (range(len(problems)) if idx is None else idx):
        off, verts = batch.vpolys(i)
        out.append((i, oracle_mod.Oracle(problems[i].with_vpolys(off, verts))))
    return out


def _check_eval(oracle_mod, problems, sigmas=(0.0, 0.05, 0.5), idx=None, seed=0):
    batch = api.Batch(problems)
    ors = _oracles(oracle_mod, batch, problems, idx)
    x0 = batch.initial_guess()
    rng = np.random.default_rng(seed)
    for s in sigmas:
        x = x0 + s * rng.normal(size=x0.size)
        cost, grad = batch.eval(x)
        dbg = batch.debug_fetch()
        for i, o in ors:
            xs = x[batch.x_off[i]:batch.x_off[i + 1]]
            f, g = o.eval(xs)
            T, q, c = o.forward(xs)
            p0, p1 = batch.piece_off[i], batch.piece_off[i + 1]
            assert np.isfinite(f)
            assert abs(cost[i] - f) <= COST_RTOL * abs(f), (i, s, cost[i], f)
            assert rel_err(grad[batch.x_off[i]:batch.x_off[i + 1]], g) <= GRAD_RTOL, (i, s)
            assert rel_err(dbg["T"][p0:p1], T) <= 1e-14
            assert rel_err(dbg["c"][18 * p0:18 * p1], c) <= 1e-11
            if o.pieces > 1:
                assert rel_err(dbg["q"][3 * p0:3 * (p1 - 1)], q) <= 1e-13
    for _, o in ors:
        o.close()
    batch.close()


# ---- integer tables through the batch API: bit-exact --------------------------------------------
def test_tables_bit_exact(oracle_mod):
    probs = [synth.make_problem(3, 6, 12, 8), api.Problem(*[getattr(synth.make_problem(4, 5, 10, 8), k) for k in ("ini", "fin", "face_off", "faces")],
                                                         dict(api.ZHANGJIAJIE, grid_res=1.3, itg_spaces=8))]
    batch = api.Batch(probs)
    for i, o in _oracles(oracle_mod, batch, probs):
        tab, ot = batch.tables(i), o.tables()
        for k in ot:
            assert np.array_equal(tab[k], ot[k]), (i, k)
        d = batch.dims(i)
        assert (d["n"], d["pieces"], d["dim_t"], d["dim_p"]) == (o.n, o.pieces, o.dim_t, o.dim_p)
    assert batch.dims(1)["pieces"] > probs[1].n_poly
    batch.close()


# ---- single evaluations vs the oracle -------------------------------------------------------------
def test_eval_parity_small_batch(oracle_mod):
    _check_eval(oracle_mod, synth.make_batch(1, 4, 8, 12, 16))


def test_eval_parity_c1_no_sfc(oracle_mod):
    _check_eval(oracle_mod, synth.config_problems("C1", 1))


def test_eval_parity_c2_track(oracle_mod):
    _check_eval(oracle_mod, synth.config_problems("C2", 1))


def test_eval_parity_c3_slice(oracle_mod):
    _check_eval(oracle_mod, synth.config_problems("C3", 16), idx=[0, 7, 15])


def test_eval_parity_c5_slice(oracle_mod):
    _check_eval(oracle_mod, synth.config_problems("C5", 8), idx=[0, 3, 7])


@pytest.mark.parametrize("K", [1, 2, 8, 15, 16, 31, 33, 48, 64, 128])
def test_eval_parity_sample_counts(oracle_mod, K):
    """Every chunking of the K+1 samples over threads (incl. K=1: both samples are end points)."""
    _check_eval(oracle_mod, synth.make_batch(20 + K, 2, 5, 10, K), sigmas=(0.3,))


def test_eval_parity_ragged_batch(oracle_mod):
    """Different piece counts, face counts, quadrature orders and parameter sets in ONE batch; M=1 (no waypoint)."""
    probs = [synth.make_problem(31, 1, 8, 4), synth.make_problem(32, 2, 6, 7), synth.make_problem(33, 9, np.array([6, 8, 30, 12, 7, 9, 50, 6, 11]), 12),
             synth.make_problem(34, 17, 12, 5, params=api.MODERNCITY), synth.make_problem(35, 3, 40, 33),
             synth.make_problem(36, 40, 10, 9)]
    _check_eval(oracle_mod, probs)


def test_eval_parity_large_shapes(oracle_mod):
    """Limits of the reference's CUDA path and beyond (cc.cuh:23-25: <= 100 pieces, <= 50 faces, <= 64 samples):
    150 pieces, 50-face polytopes, 129 samples per piece."""
    probs = [synth.make_problem(91, 150, 8, 4), synth.make_problem(92, 4, 50, 128), synth.make_problem(93, 100, np.full(100, 9), 6)]
    _check_eval(oracle_mod, probs, sigmas=(0.0, 0.3))


def test_capacity_limit_is_reported():
    """The band system of one trajectory must fit one SM's shared memory: far beyond that, create() says so."""
    p = synth.make_problem(94, 500, 6, 2)
    with pytest.raises(api.FrxError) as e:
        api.Batch([p])
    assert e.value.status == api.FRX_ERR_CAPACITY


def test_batches_of_different_shapes_coexist(oracle_mod):
    """Distinct batches are independent (frx.h): creating a SMALLER batch must not break a live larger one (the
    per-function dynamic shared-memory limit is process-global and may only grow)."""
    big = [synth.make_problem(95, 120, np.full(120, 30), 64)]
    small = [synth.make_problem(96, 3, 6, 4)]
    bb = api.Batch(big)
    xb = bb.initial_guess() + 0.1 * np.random.default_rng(5).normal(size=bb.total_x)
    c0, g0 = bb.eval(xb)
    bs = api.Batch(small)
    cs, _ = bs.eval(bs.initial_guess())
    c1, g1 = bb.eval(xb)  # the larger batch, after the smaller one configured the kernels
    assert np.isfinite(cs).all() and np.isfinite(c1).all()
    assert np.array_equal(c0, c1) and np.array_equal(g0, g1)
    off, verts = bb.vpolys(0)
    o = oracle_mod.Oracle(big[0].with_vpolys(off, verts))
    f, g = o.eval(xb)
    assert abs(c1[0] - f) <= COST_RTOL * abs(f) and rel_err(g1, g) <= GRAD_RTOL
    o.close()
    bs.close()
    bb.close()


def test_entry_points_leave_the_callers_current_device_alone():
    """Every C-ABI entry point runs on its handle's device and restores the caller's current device (a planner with several
    GPUs, or torch in the same process, keeps its own).  Needs two GPUs to be observable."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rt = C.CDLL("libcudart.so.12")
    cur = C.c_int(-1)
    rt.cudaSetDevice(0)
    b = api.Batch(synth.make_batch(420, 2, 4, 8, 8), device=1)
    rt.cudaGetDevice(C.byref(cur))
    assert cur.value == 0
    c, g = b.eval(b.initial_guess())
    r = b.optimize(1e-3)
    rt.cudaGetDevice(C.byref(cur))
    assert cur.value == 0 and np.isfinite(c).all() and np.isfinite(r["final_cost"]).all()
    b.close()
    rt.cudaGetDevice(C.byref(cur))
    assert cur.value == 0


def test_solvers_of_different_shapes_coexist():
    """The update kernel's dynamic shared-memory limit (ring of the streamed history) is per function and process-global too:
    a solver created later for SMALLER trajectories must not lower it under a live larger one.  Solves are deterministic, so
    the larger batch must reproduce its own result bit for bit afterwards."""
    big = api.Batch(synth.make_batch(400, 3, 40, 12, 16))
    r0 = big.optimize(1e-4)
    small = api.Batch(synth.make_batch(410, 2, 3, 8, 8))
    rs = small.optimize(1e-4)
    assert np.isfinite(rs["final_cost"]).all()
    r1 = big.optimize(1e-4)
    for k in ("x", "status", "iterations", "evaluations", "final_cost"):
        assert np.array_equal(r0[k], r1[k]), k
    small.close()
    big.close()


def test_device_vpolytope_extraction_matches_host_path(oracle_mod):
    """frx_batch_create extracts the V-polytopes of all corridors in one kernel launch (one warp per polytope);
    the result is bit-identical — vertices AND order — to the host enumeration (same algorithm, geometry_core.hpp), equals the
    oracle's brute-force vertex sets, and infeasible corridors are still rejected (SE3GCOPTER::setup returning false)."""
    probs = [synth.make_problem(81, 9, np.array([6, 8, 30, 12, 7, 9, 50, 6, 11]), 6), synth.make_problem(82, 1, 8, 4),
             synth.config_problems("C2", 1)[0], synth.make_problem(83, 12, 12, 8)]
    b = api.Batch(probs)
    for i, p in enumerate(probs):
        off_d, v_d = b.vpolys(i)
        off_h, v_h = api.host_vpolys(p)
        assert np.array_equal(off_d, off_h) and np.array_equal(v_d, v_h), i
        faces = p.faces.reshape(-1, 6).copy()
        faces[:, :3] /= np.linalg.norm(faces[:, :3], axis=1, keepdims=True)
        for k in range(len(off_d) - 1):
            a = v_d[3 * off_d[k]:3 * off_d[k + 1]].reshape(-1, 3).copy()
            a[1:] += a[0]
            j = k // 2
            f = faces[p.face_off[j]:p.face_off[j + 1]] if k % 2 == 0 else faces[p.face_off[j]:p.face_off[j + 2]]
            bf = oracle_mod.brute_force_vertices(f)
            assert len(bf) == len(a) and np.abs(a[:, None] - bf[None]).sum(-1).min(1).max() < 1e-8, (i, k)
    b.close()
    bad = synth.make_problem(84, 3, 8, 8)
    faces = bad.faces.reshape(-1, 6).copy()
    faces[bad.face_off[1]:bad.face_off[2], 3:] += 1000.0  # intersection 0^1 is empty
    with pytest.raises(api.FrxError) as e:
        api.Batch([synth.make_problem(85, 3, 8, 8), api.Problem(bad.ini, bad.fin, bad.face_off, faces, bad.params)])
    assert e.value.status == api.FRX_ERR_INFEASIBLE_GEOMETRY


def test_device_initial_guess_is_bit_identical_to_host_path(monkeypatch):
    """the M - 1 per-waypoint solves of backwardP (se3gcopter_cpu.hpp:777-813) run on the device, one thread
    each, with the same code as the host path (waypoint_nls.hpp, no FMA contraction): identical bits — and, through the golden
    fixtures (test_cuda_matches_golden_vectors), identical to the reference's own setInitial + backwardT + backwardP."""
    probs = [synth.make_problem(86, 9, np.array([6, 8, 30, 12, 7, 9, 50, 6, 11]), 6), synth.make_problem(87, 1, 8, 4),
             synth.config_problems("C2", 1)[0], synth.make_problem(88, 40, 10, 9, params=api.MODERNCITY)]
    probs.append(api.Problem(probs[0].ini, probs[0].fin, probs[0].face_off, probs[0].faces, dict(probs[0].params, grid_res=1.1, c2_diffeo=0)))
    b = api.Batch(probs)
    x_dev = b.initial_guess()
    monkeypatch.setenv("FRX_HOST_INIT", "1")
    x_host = b.initial_guess()
    assert np.isfinite(x_dev).all() and np.array_equal(x_dev, x_host)
    b.close()


def test_eval_parity_time_maps(oracle_mod):
    """Fixed total time (rho = 0) with the C2 and the exp diffeomorphism; soft time with exp."""
    p = synth.make_problem(41, 6, 10, 8)
    probs = [api.Problem(p.ini, p.fin, p.face_off, p.faces, dict(p.params, rho=rho, total_t=tt, c2_diffeo=c2))
             for rho, tt, c2 in ((0.0, 5.0, 1), (0.0, 5.0, 0), (800.0, 0.0, 0), (1000.0, 0.0, 1))]
    _check_eval(oracle_mod, probs, sigmas=(0.0, 0.2))


def test_eval_parity_interval_splitting(oracle_mod):
    """gridRes < inf: several pieces per polytope (splitToFineT / mergeToCoarseGradT / idxVs 2i and 2i+1)."""
    p = synth.make_problem(42, 5, 12, 8)
    probs = [api.Problem(p.ini, p.fin, p.face_off, p.faces, dict(p.params, grid_res=g)) for g in (1.1, 2.5)]
    _check_eval(oracle_mod, probs, sigmas=(0.0, 0.2))


def test_nonzero_boundary_states(oracle_mod):
    p = synth.make_problem(43, 6, 12, 8)
    ini = p.ini.copy()
    fin = p.fin.copy()
    ini[3:6] = [3.0, -1.0, 0.5]
    ini[6:9] = [0.5, 0.2, -0.3]
    fin[3:6] = [30.0, 0.0, 0.0]  # > vMax: the reference's clamp is a dead store — must NOT be applied
    _check_eval(oracle_mod, [api.Problem(ini, fin, p.face_off, p.faces, p.params)], sigmas=(0.0, 0.3))


# ---- golden fixtures ----------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names())
def test_cuda_matches_golden_vectors(name):
    p, z = load_golden(name)
    batch = api.Batch([p])
    d = batch.dims(0)
    assert [d["n"], d["pieces"], d["dim_t"], d["dim_p"]] == z["dims"].tolist()
    tab = batch.tables(0)
    for k, v in tab.items():
        assert np.array_equal(v, z["tab_" + k])
    assert np.array_equal(batch.initial_guess(), z["x0"])
    for i, x in enumerate(z["xs"]):
        cost, grad = batch.eval(x)
        # ref_* = outputs of the reference's own code (verbatim se3gcopter_cpu.hpp over the Eigen-API shim)
        assert abs(cost[0] - z["ref_cost"][i]) <= COST_RTOL * abs(z["ref_cost"][i])
        assert rel_err(grad, z["ref_grad"][i]) <= GRAD_RTOL
        assert abs(cost[0] - z["cost"][i]) <= COST_RTOL * abs(z["cost"][i])
        assert rel_err(grad, z["grad"][i]) <= GRAD_RTOL
        dbg = batch.debug_fetch()
        assert rel_err(dbg["c"], z["c"][i]) <= 1e-11
    batch.close()


# ---- solves -------------------------------------------------------------------------------------------
# The reference stops when the relative decrease over 3 iterations drops below delta (lbfgs.hpp:1327-1345) — on
# these penalty landscapes that fires on plateaus, and WHERE it fires is chaotic in the last bits of the
# arithmetic: the oracle itself, restarted from x0 perturbed by k ulps, spreads by ~1e-3 relative at delta = 1e-5
# and still by ~1e-5 at delta = 1e-9; only at delta <= 1e-12 do (most) runs agree to ~1e-9.  Parity of a solve is
# therefore stated against the oracle's own ulp-perturbation band (DESIGN.md "Numerics"):
#   converged (delta = 1e-13):   cost inside [lo, hi] of the band widened by 1e-6 relative,
#   reference delta (1e-5):      |cost - band midpoint| <= 3 x max(band width, 0.2 %).
def _oracle_solve_from(oracle_mod, o, x_start, tol):
    """Oracle objective minimised from an arbitrary start: by the verbatim reference lbfgs.hpp when oracle/_ref is
    present, else by the (bit-identical) restated L-BFGS.  Returns the cost AT the returned x."""
    R = oracle_mod.ref_lbfgs()
    L = oracle_mod.lib()
    x = np.ascontiguousarray(x_start, dtype=np.float64).copy()
    f = C.c_double()
    if R is not None:
        cb = oracle_mod.EVAL_T(lambda inst, xx, g, n: L.fro_eval(o.h, xx, g, n))
        R.ref_lbfgs_optimize(len(x), x.ctypes.data_as(C.POINTER(C.c_double)), C.byref(f), cb, None, 128, 3, tol, 1e-16, 1e-32, 0)
    else:
        L.fro_optimize_from(o.h, x.ctypes.data_as(C.POINTER(C.c_double)), tol)
    return o.eval(x)[0]


def _oracle_band(oracle_mod, o, tol, runs=5):
    x0 = o.initial_guess()
    vals = [_oracle_solve_from(oracle_mod, o, x0 * (1.0 + k * 2.2e-16), tol) for k in range(runs)]
    return min(vals), max(vals)


def test_reference_lbfgs_drives_cuda_evaluator(oracle_mod):
    """The lbfgs_evaluate_t boundary: the VERBATIM reference lbfgs_optimize minimises through frx_eval (CUDA)."""
    R = oracle_mod.ref_lbfgs()
    if R is None:
        pytest.skip("oracle/_ref (verbatim lbfgs.hpp) not available on this box")
    p, z = load_golden("sfc_small")
    batch = api.Batch([p])
    fn, inst = batch.eval_callback()
    x = z["x0"].copy()
    f = C.c_double()
    cb = C.cast(fn, oracle_mod.EVAL_T)
    tol = 1e-13
    R.ref_lbfgs_optimize(len(x), x.ctypes.data_as(C.POINTER(C.c_double)), C.byref(f), cb, inst, 128, 3, tol, 1e-16, 1e-32, 0)
    o = oracle_mod.Oracle(p)
    lo, hi = _oracle_band(oracle_mod, o, tol)
    fx = o.eval(x)[0]
    assert lo * (1 - 1e-6) <= fx <= hi * (1 + 1e-6), (fx, lo, hi)
    batch.close()


def test_solver_converged_cost_matches_oracle(oracle_mod):
    """Batched device L-BFGS run to convergence reaches the oracle's minimum: cost within 1e-6 relative."""
    probs = synth.make_batch(50, 6, 6, 10, 8)
    batch = api.Batch(probs)
    tol = 1e-13
    r = batch.optimize(tol)
    x0 = batch.initial_guess()
    n_converged = 0
    # statuses of a run that went as far as the arithmetic allows: the delta / gradient stops, or a line search that can
    # no longer resolve a decrease below the rounding noise of f (delta = 1e-13 is 7 orders tighter than the reference's)
    rounding_exits = (-1005, -1001, -1003, -1007, -1008)  # MAXIMUMLINESEARCH, INCREASEGRADIENT, WIDTHTOOSMALL, MINIMUMSTEP, ROUNDING_ERROR
    for i, o in _oracles(oracle_mod, batch, probs):
        p0, p1 = batch.x_off[i], batch.x_off[i + 1]
        f, _ = o.eval(r["x"][p0:p1])  # oracle cost at the GPU's final iterate
        lo, hi = _oracle_band(oracle_mod, o, tol)
        # north_star: converged costs within 1e-6 relative of the reference's — for EVERY trajectory, whatever the exit
        assert lo * (1 - 1e-6) <= f <= hi * (1 + 1e-6), (i, int(r["status"][i]), f, lo, hi)
        if r["status"][i] >= 0:
            assert abs(f - r["final_cost"][i]) <= 1e-10 * f
            n_converged += 1
        else:
            # A line search can fail on some rounding paths (the reference's own lbfgs_optimize, fed these evaluations,
            # returns the same error codes; SE3GCOPTER::optimize swallows them, CPU.hpp:1249): the iterate returned is
            # then the last accepted one.  Bit-identical paths are the reference-order tests.
            assert int(r["status"][i]) in rounding_exits, (i, r["status"][i])
            assert f < o.eval(x0[p0:p1])[0], (i, r["status"][i])
        # output trajectory: durations / coefficients / jerk cost consistent with the oracle at that x
        T, q, c = o.forward(r["x"][p0:p1])
        a0, a1 = batch.piece_off[i], batch.piece_off[i + 1]
        assert rel_err(r["durations"][a0:a1], T) <= 1e-13
        co = r["coeffs"][18 * a0:18 * a1].reshape(-1, 6, 3)  # [piece][col][dim], highest order first
        assert rel_err(co[:, ::-1, :].reshape(-1), c) <= 1e-10
        assert np.isfinite(r["jerk_cost"][i]) and 0 < r["jerk_cost"][i] < f
    assert n_converged >= 1, r["status"]
    bi, bc = batch.argmin()
    assert bi == int(np.argmin(r["final_cost"])) and bc == r["final_cost"].min()
    batch.close()


def test_vector_free_direction_matches_two_loop_recursion(monkeypatch):
    """The vector-free direction update (Gram-matrix recursions as blocked right-looking solves) against the sequential
    two-loop recursion of lbfgs.hpp:1359-1411: with FRX_LBFGS_CHECK=1 the update kernel computes BOTH at every iteration
    of a full solve (history growing through every block size up to the full 128 pairs) and records the largest deviation."""
    monkeypatch.setenv("FRX_LBFGS_CHECK", "1")
    probs = synth.make_batch(4, 8, 10, 12, 31)
    b = api.Batch(probs)
    r = b.optimize(1e-9)
    assert (r["iterations"] > 200).any(), r["iterations"]  # the history filled up
    assert 0.0 <= r["direction_check"] <= 1e-9, r["direction_check"]
    b.close()


def test_solver_reference_tolerance_within_sensitivity_band(oracle_mod):
    """At the reference's own delta = 1e-5 the GPU result lies inside the oracle's ulp-perturbation band.  The band is
    taken over 24 perturbed oracle solves: the delta test stops on plateaus, and the plateau a solve stops on depends on
    its rounding path — on the second trajectory here the ORACLE's own final costs under 1-ulp perturbations of the start
    point spread over 4678 ... 4768 (2 %), with most runs near the lower end, so a 6-run band misses the upper cluster."""
    probs = synth.make_batch(60, 4, 8, 12, 16)
    batch = api.Batch(probs)
    tol = 1e-5
    r = batch.optimize(tol)
    assert (r["status"] >= 0).all()
    for i, o in _oracles(oracle_mod, batch, probs):
        lo, hi = _oracle_band(oracle_mod, o, tol, runs=24)
        mid = 0.5 * (lo + hi)
        band = max(hi - lo, 2e-3 * mid)  # the oracle's own spread under ulp perturbations (floor 0.2 %)
        assert abs(r["final_cost"][i] - mid) <= 1.5 * band, (i, r["final_cost"][i], lo, hi)
    batch.close()


# ---- reference-order mode: bit-identical to the reference CPU code ---------------------------------------------
@pytest.mark.parametrize("name", golden_names())
def test_reference_order_mode_is_bit_identical_to_reference(name):
    """frx_batch_set_reference_order: the CUDA evaluation and the CUDA batched L-BFGS reproduce the golden vectors —
    which are the outputs of the reference's own code (verbatim se3gcopter_cpu.hpp over the Eigen-API shim) — BIT FOR
    BIT: cost, gradient, coefficients, and the whole solve at the reference's stopping tolerance (status, iteration and
    evaluation counts, final x, final cost, durations)."""
    p, z = load_golden(name)
    batch = api.Batch([p])
    batch.set_reference_order(True)
    for i, x in enumerate(z["xs"]):
        cost, grad = batch.eval(x)
        assert cost[0] == z["ref_cost"][i] == z["cost"][i]
        assert np.array_equal(grad, z["ref_grad"][i])
        dbg = batch.debug_fetch()
        assert np.array_equal(dbg["c"], z["c"][i]) and np.array_equal(dbg["T"], z["T"][i])
    r = batch.optimize(float(z["opt_rel_tol"]))
    assert r["status"][0] == int(z["opt_status"])
    assert r["iterations"][0] == int(z["opt_iterations"]) and r["evaluations"][0] == int(z["opt_evaluations"])
    assert np.array_equal(r["x"], z["opt_x"])
    assert r["final_cost"][0] == float(z["opt_final_cost"])
    assert np.array_equal(r["durations"], z["ref_opt_durations"])
    assert np.array_equal(r["coeffs"], z["ref_opt_coeffs"])
    assert r["jerk_cost"][0] == float(z["ref_opt_jerk_cost"])
    batch.close()


def test_reference_order_mode_batch_matches_oracle_bitwise(oracle_mod):
    """A ragged batch (different M / F / K / parameter sets) in reference-order mode: every trajectory's solve equals the
    oracle's bit for bit; the fast mode stays within the stated tolerance of it."""
    probs = [synth.make_problem(81, 3, 8, 6), synth.make_problem(82, 7, np.array([6, 9, 12, 8, 7, 10, 6]), 9),
             synth.make_problem(83, 5, 10, 12, params=api.MODERNCITY), synth.make_problem(84, 6, 12, 8, gates=(1, 3))]
    batch = api.Batch(probs)
    x = batch.initial_guess() + 0.3 * np.random.default_rng(2).normal(size=batch.total_x)
    c_fast, g_fast = batch.eval(x)
    batch.set_reference_order(True)
    c_ex, g_ex = batch.eval(x)
    r = batch.optimize(1e-5)
    for i, o in _oracles(oracle_mod, batch, probs):
        s = slice(batch.x_off[i], batch.x_off[i + 1])
        f, g = o.eval(x[s])
        assert c_ex[i] == f and np.array_equal(g_ex[s], g)
        assert abs(c_fast[i] - f) <= COST_RTOL * abs(f) and rel_err(g_fast[s], g) <= GRAD_RTOL
        ro = o.optimize(1e-5)
        assert r["status"][i] == ro["status"] and r["iterations"][i] == ro["iterations"] and r["evaluations"][i] == ro["evaluations"]
        assert np.array_equal(r["x"][s], ro["x"]) and r["final_cost"][i] == ro["final_cost"]
        assert r["jerk_cost"][i] == ro["jerk_cost"]
    batch.set_reference_order(False)
    c2, g2 = batch.eval(x)
    assert np.array_equal(c2, c_fast) and np.array_equal(g2, g_fast)
    batch.close()


def test_solver_statuses_and_masking(oracle_mod):
    """Trajectories of very different difficulty in one batch: early finishers are masked, all report a status."""
    probs = [synth.make_problem(70, 1, 8, 8), synth.make_problem(71, 3, 8, 8), synth.make_problem(72, 12, 12, 16)]
    batch = api.Batch(probs)
    r = batch.optimize(1e-5)
    # trajectory 0 has ONE free variable whose gradient at the optimum is the difference of terms 1e8 times larger (the
    # oracle itself needs 45 iterations / 95 evaluations on it): its line search may end on a rounding exit instead of the
    # delta test — both are "went as far as the arithmetic allows"; the cost checks below hold either way
    assert all(st >= 0 or st in (-1005, -1001, -1003, -1007, -1008) for st in r["status"]), r["status"]
    assert (r["status"][1:] >= 0).all(), r["status"]
    assert r["evaluations"][0] < r["evaluations"][2]
    assert r["ticks"] >= r["evaluations"].max()
    for i, o in _oracles(oracle_mod, batch, probs):
        f, _ = o.eval(r["x"][batch.x_off[i]:batch.x_off[i + 1]])
        assert abs(f - r["final_cost"][i]) <= 1e-10 * abs(f)  # final_cost is the objective AT the returned iterate, whatever the exit
        if i == 0:
            assert abs(f - o.optimize(1e-5)["final_cost"]) <= 1e-6 * f  # one variable: the minimum itself is well defined
    batch.close()


def test_solver_slots_do_not_change_results():
    """Queue-fed solver slots (frx_batch_set_solver_slots): every trajectory's solve is independent of what else is in flight,
    so the iterates, statuses and evaluation counts are bit-identical whether 24 ragged trajectories get a slot each or share
    1, 5 or 23 slots — in both builds of the device code.  Evaluating afterwards sees no stale slot list."""
    probs = [synth.make_problem(300 + i, 2 + (5 * i) % 11, 6 + i % 5, 8 + 4 * (i % 3)) for i in range(24)]
    for exact in (0, 1):
        batch = api.Batch(probs)
        batch.set_reference_order(exact)
        ref = batch.optimize(1e-5)
        assert (ref["evaluations"] > 0).all()
        for slots in (1, 5, 23, 24, 100):
            batch.set_solver_slots(slots)
            r = batch.optimize(1e-5)
            for k in ("x", "status", "iterations", "evaluations", "final_cost", "durations", "coeffs", "jerk_cost"):
                assert np.array_equal(ref[k], r[k]), (exact, slots, k)
            if slots < 24:  # fewer slots than trajectories: waves follow each other
                assert r["ticks"] >= ref["ticks"] and (slots > 5 or r["ticks"] > ref["ticks"])
        x = batch.initial_guess()
        c1, g1 = batch.eval(x)
        batch.set_solver_slots(0)
        c2, g2 = batch.eval(x)
        assert np.array_equal(c1, c2) and np.array_equal(g1, g2) and np.isfinite(c1).all()
        batch.close()


# ---- properties at BASELINE.json's full sizes (C4: 1024 x 64 x 33 x 12) ---------------------------
@pytest.fixture(scope="module")
def c4_batch():
    probs = synth.config_problems("C4", 1024)
    b = api.Batch(probs)
    yield b, probs
    b.close()


def test_c4_determinism_and_oracle_spot_checks(oracle_mod, c4_batch):
    batch, probs = c4_batch
    assert batch.count == 1024 and batch.dims(0)["pieces"] == 64
    x = batch.initial_guess() + 0.5 * np.random.default_rng(5).normal(size=batch.total_x)
    c1, g1 = batch.eval(x)
    c2, g2 = batch.eval(x)
    assert np.array_equal(c1, c2) and np.array_equal(g1, g2)  # fixed reduction order: bitwise reproducible
    assert np.isfinite(c1).all() and np.isfinite(g1).all()
    for i, o in _oracles(oracle_mod, batch, probs, idx=[0, 511, 1023]):
        f, g = o.eval(x[batch.x_off[i]:batch.x_off[i + 1]])
        assert abs(c1[i] - f) <= COST_RTOL * abs(f)
        assert rel_err(g1[batch.x_off[i]:batch.x_off[i + 1]], g) <= GRAD_RTOL
        o.close()


def test_c4_batch_order_independence(c4_batch):
    """A trajectory's result does not depend on its position in the batch or on its neighbours (bitwise)."""
    batch, probs = c4_batch
    x = batch.initial_guess()
    c_full, g_full = batch.eval(x)
    perm = [1023, 0, 700, 17]
    sub = api.Batch([probs[i] for i in perm])
    xs = np.concatenate([x[batch.x_off[i]:batch.x_off[i + 1]] for i in perm])
    c_sub, g_sub = sub.eval(xs)
    for k, i in enumerate(perm):
        assert c_sub[k] == c_full[i]
        assert np.array_equal(g_sub[sub.x_off[k]:sub.x_off[k + 1]], g_full[batch.x_off[i]:batch.x_off[i + 1]])
    sub.close()


def test_c4_fast_mode_agrees_with_reference_order_mode_on_every_trajectory(c4_batch):
    """Fast mode (reduced system by cyclic reduction, adjoint-vector penalties, tree reductions) against reference-order mode
    (bit-identical to the reference's own code) on ALL 1024 trajectories of C4, at a perturbed point with every penalty
    family active: the stated fp64 tolerance of the fast path, not just spot checks."""
    batch, _ = c4_batch
    x = batch.initial_guess() + 0.3 * np.random.default_rng(9).normal(size=batch.total_x)
    c_fast, g_fast = batch.eval(x)
    batch.set_reference_order(1)
    c_ref, g_ref = batch.eval(x)
    batch.set_reference_order(0)
    rc = np.abs(c_fast - c_ref) / np.abs(c_ref)
    rg = np.array([rel_err(g_fast[batch.x_off[i]:batch.x_off[i + 1]], g_ref[batch.x_off[i]:batch.x_off[i + 1]]) for i in range(batch.count)])
    assert np.median(rc) <= 1e-13 and rc.max() <= 5e-11, (np.median(rc), rc.max())
    assert np.median(rg) <= 1e-12 and rg.max() <= GRAD_RTOL, (np.median(rg), rg.max())


def test_c4_gradient_is_the_derivative_of_the_cost(c4_batch):
    """Directional finite differences of the CUDA cost against the CUDA gradient, all 1024 trajectories at once."""
    batch, _ = c4_batch
    rng = np.random.default_rng(9)
    x = batch.initial_guess() + 0.3 * rng.normal(size=batch.total_x)
    d = rng.normal(size=batch.total_x)
    for i in range(batch.count):
        s = slice(batch.x_off[i], batch.x_off[i + 1])
        d[s] /= np.linalg.norm(d[s])
    # central differences: rounding noise eps_cost * |cost| / h against truncation ~ h^2: the cost carries ~1e-11 relative
    # rounding in the worst trajectories (cyclic reduction of the MINCO system), so h = 1e-5
    h = 1e-5
    c0, g0 = batch.eval(x)
    cp, _ = batch.eval(x + h * d)
    cm, _ = batch.eval(x - h * d)
    fd = (cp - cm) / (2 * h)
    gd = np.array([g0[batch.x_off[i]:batch.x_off[i + 1]] @ d[batch.x_off[i]:batch.x_off[i + 1]] for i in range(batch.count)])
    err = np.abs(fd - gd) / np.maximum(np.abs(fd), np.abs(gd))
    assert np.median(err) < 1e-7 and err.max() < 1e-4


def test_pinned_host_buffer_api_matches_staged_api(c4_batch):
    batch, _ = c4_batch
    x = batch.initial_guess() + 0.1 * np.random.default_rng(3).normal(size=batch.total_x)
    c_h, g_h = batch.eval(x)
    xp, cp, gp = batch.host_buffers()
    xp[:] = x
    batch.eval_pinned()
    assert np.array_equal(cp, c_h) and np.array_equal(gp, g_h)


def test_device_pointer_api_matches_host_api(c4_batch):
    import torch

    batch, _ = c4_batch
    x = batch.initial_guess()
    c_h, g_h = batch.eval(x)
    xd = torch.from_numpy(x).cuda()
    gd = torch.zeros_like(xd)
    cd = torch.zeros(batch.count, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream()
    batch.eval_device(xd.data_ptr(), cd.data_ptr(), gd.data_ptr(), st.cuda_stream)
    st.synchronize()
    assert np.array_equal(cd.cpu().numpy(), c_h) and np.array_equal(gd.cpu().numpy(), g_h)
    ms = batch.profile_eval(xd.data_ptr(), cd.data_ptr(), gd.data_ptr(), st.cuda_stream, 2)
    assert (ms > 0).all()
