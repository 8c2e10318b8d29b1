"""Generates the committed golden fixtures tests/golden/*.npz.

Run in the authoring container (where /root/reference exists):  python tests/golden/make_golden.py
Each fixture holds one small problem (faces, V-polytopes, parameters), evaluation points, and the outputs of
  * the CPU oracle (oracle/gcopter_oracle.hpp): cost, gradient, T, q, c, index tables, initial guess,
  * the oracle's objective driven by the VERBATIM reference lbfgs.hpp (oracle/_ref/libref_lbfgs.so):
    final x / cost / status — and the restated L-BFGS must reproduce them bit for bit,
  * the reference's OWN code — verbatim se3gcopter_cpu.hpp + trajectory.hpp + lbfgs.hpp compiled over the Eigen-API shim
    (oracle/_ref/libfrx_ref.so): cost / gradient at the same points ("ref_cost", "ref_grad") and its complete optimize()
    result ("ref_opt_*") — asserted BIT-IDENTICAL to the oracle's at generation time.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fast_racing_b200 import api, synth  # noqa: E402
from oracle import binding as ob  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (problem factory, perturbation sigma)
    "c1_small": (lambda: synth.make_problem(101, 8, 10, 12, chi=(0.0, 1.0e4, 1.0e4, 1.0e4), gates=(1, 3, 5, 7)), 0.05),
    "sfc_small": (lambda: synth.make_problem(102, 6, 12, 8), 0.05),
    "ragged_faces": (lambda: synth.make_problem(103, 5, np.array([8, 14, 6, 20, 9]), 10, gates=(2,)), 0.1),
    "moderncity_small": (lambda: synth.make_problem(104, 7, 12, 16, params=api.MODERNCITY), 0.05),
    # "*_refgeo": the V-polytopes are the REFERENCE'S OWN (verbatim geoutils.hpp + sdlp.hpp + quickhull.hpp through
    # SE3GCOPTER::setup, oracle/_ref/libfrx_ref_full.so), i.e. the vertex order that defines the xi variables is the
    # reference's; every ref_* output below then comes from the fully verbatim reference.
    "sfc_small_refgeo": (lambda: synth.make_problem(102, 6, 12, 8), 0.05),
    "ragged_faces_refgeo": (lambda: synth.make_problem(103, 5, np.array([8, 14, 6, 20, 9]), 10, gates=(2,)), 0.1),
}


def main():
    R = ob.ref_lbfgs()
    L = ob.lib()
    only = set(sys.argv[1:])
    for name, (factory, sigma) in CASES.items():
        if only and name not in only:
            continue
        p = factory()
        refgeo = name.endswith("_refgeo")
        if refgeo:
            full = ob.RefGcopterFull(p)  # the reference's own setup(): LP interior point + dual quickhull + first-seen dedup
            off, verts = full.vpolys()
            po, pvv = api.host_vpolys(p)  # the product's enumeration: same vertex SETS, its own order
            assert np.array_equal(np.diff(off), np.diff(po))
            for k in range(len(off) - 1):
                a = verts[3 * off[k]:3 * off[k + 1]].reshape(-1, 3).copy()
                b = pvv[3 * po[k]:3 * po[k + 1]].reshape(-1, 3).copy()
                a[1:] += a[0]
                b[1:] += b[0]
                assert np.abs(a[:, None] - b[None]).sum(-1).min(1).max() < 1e-9, (name, k)
        else:
            off, verts = api.host_vpolys(p)
        pv = p.with_vpolys(off, verts)
        o = ob.Oracle(pv)
        x0 = o.initial_guess()
        rng = np.random.default_rng(abs(hash(name)) % (2**31))
        rng = np.random.default_rng(sum(map(ord, name)))
        xs = np.stack([x0, x0 + sigma * rng.normal(size=x0.size), x0 + 3 * sigma * rng.normal(size=x0.size)])
        costs, grads, Ts, qs, cs = [], [], [], [], []
        for x in xs:
            f, g = o.eval(x)
            T, q, c = o.forward(x)
            costs.append(f), grads.append(g), Ts.append(T), qs.append(q), cs.append(c)
        tab = o.tables()
        out = dict(ini=p.ini, fin=p.fin, face_off=p.face_off, faces=p.faces, vert_off=off, verts=verts,
                   params_keys=np.array(sorted(k for k in p.params if k != "chi")),
                   params_vals=np.array([float(p.params[k]) for k in sorted(p.params) if k != "chi"]),
                   chi=np.array(p.params["chi"]), xs=xs, cost=np.array(costs), grad=np.stack(grads), T=np.stack(Ts),
                   q=np.stack(qs), c=np.stack(cs), x0=x0, **{f"tab_{k}": v for k, v in tab.items()},
                   dims=np.array([o.n, o.pieces, o.dim_t, o.dim_p]))
        tol = p.params.get("rel_tol", 1e-5)
        r = o.optimize(tol)
        out.update(opt_x=r["x"], opt_final_cost=r["final_cost"], opt_jerk_cost=r["jerk_cost"], opt_status=r["status"],
                   opt_iterations=r["iterations"], opt_evaluations=r["evaluations"], opt_durations=r["durations"],
                   opt_coeffs=r["coeffs"], opt_rel_tol=tol)
        if ob.RefGcopter.available():
            # the reference's OWN code (verbatim se3gcopter_cpu.hpp over the Eigen-API shim) on the same inputs
            rg = full if refgeo else ob.RefGcopter(pv)
            rc, rgr = zip(*[rg.eval(x) for x in xs])
            assert np.array_equal(np.array(rc), np.array(costs)) and np.array_equal(np.stack(rgr), np.stack(grads)), \
                f"{name}: oracle deviates from the verbatim reference source"
            assert np.array_equal(rg.initial_guess(), x0)
            ro = rg.optimize(tol)
            assert ro["jerk_cost"] == r["jerk_cost"] and np.array_equal(ro["durations"], r["durations"])
            out.update(ref_cost=np.array(rc), ref_grad=np.stack(rgr), ref_opt_jerk_cost=ro["jerk_cost"], ref_opt_durations=ro["durations"],
                       ref_opt_coeffs=ro["coeffs"])
            rg.close()
        if R is not None:
            xr = x0.copy()
            fr = C.c_double()
            cb = ob.EVAL_T(lambda inst, x, g, n: L.fro_eval(o.h, x, g, n))
            st = R.ref_lbfgs_optimize(o.n, xr.ctypes.data_as(C.POINTER(C.c_double)), C.byref(fr), cb, None, 128, 3, tol, 1e-16, 1e-32, 0)
            assert st == r["status"] and fr.value == r["final_cost"] and np.array_equal(xr, r["x"]), \
                f"{name}: restated L-BFGS deviates from the verbatim reference lbfgs.hpp"
            out.update(reflbfgs_x=xr, reflbfgs_cost=fr.value, reflbfgs_status=st)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, "n", o.n, "cost", costs, "opt", r["final_cost"], r["status"], r["iterations"])
        o.close()


if __name__ == "__main__":
    main()
