"""Ad-hoc GPU shake-out: eval parity vs oracle, solver parity, rough timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_racing_b200 import api, synth
from oracle import binding as ob

def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

def check(problems, tag, perturb=0.05, solve_tol=None):
    t = time.time(); batch = api.Batch(problems); print(tag, "create", round(time.time() - t, 3), "s", flush=True)
    oracles = []
    for i, p in enumerate(problems[:4]):
        off, verts = batch.vpolys(i)
        oracles.append(ob.Oracle(p.with_vpolys(off, verts)))
    x0 = batch.initial_guess()
    rng = np.random.default_rng(7)
    for xs, name in ((x0, "x0"), (x0 + perturb * rng.normal(size=x0.size), "x0+noise")):
        cost, grad = batch.eval(xs)
        dbg = batch.debug_fetch()
        for i, o in enumerate(oracles):
            xi = xs[batch.x_off[i]:batch.x_off[i + 1]]
            f, g = o.eval(xi)
            T, q, c = o.forward(xi)
            pc, pgT, pgC = o.penalty_only()
            p0, p1 = batch.piece_off[i], batch.piece_off[i + 1]
            print(f"  {tag} {name} traj{i}: cost rel {abs(cost[i]-f)/abs(f):.2e} grad rel {rel(grad[batch.x_off[i]:batch.x_off[i+1]], g):.2e}"
                  f" | T {rel(dbg['T'][p0:p1], T):.1e} c {rel(dbg['c'][18*p0:18*p1], c):.1e} pen {abs(dbg['pen'][p0:p1].sum()-pc)/max(abs(pc),1e-300):.1e}", flush=True)
    # timing of device-resident evals
    import torch
    xd = torch.from_numpy(x0).cuda(); gd = torch.zeros_like(xd); cd = torch.zeros(batch.count, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3): batch.eval_device(xd.data_ptr(), cd.data_ptr(), gd.data_ptr(), st)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps): batch.eval_device(xd.data_ptr(), cd.data_ptr(), gd.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"  {tag} device eval: {ms*1e3:.1f} us / batch-eval -> {batch.count/ms*1e3:.3e} evals/s", flush=True)
    t = time.time(); cost, grad = batch.eval(x0); print(f"  {tag} host-buffer eval: {(time.time()-t)*1e3:.2f} ms", flush=True)
    if solve_tol:
        t = time.time(); r = batch.optimize(solve_tol); dt = time.time() - t
        print(f"  {tag} optimize: wall {dt:.3f}s device {r['device_ms']:.1f} ms ticks {r['ticks']} iters {r['iterations'][:4]} evals {r['evaluations'][:4]} status {r['status'][:4]}", flush=True)
        for i, o in enumerate(oracles[:2]):
            t = time.time(); ro = o.optimize(solve_tol); dto = time.time() - t
            print(f"    traj{i}: gpu final {r['final_cost'][i]:.9e} jerk {r['jerk_cost'][i]:.6e} | oracle final {ro['final_cost']:.9e} jerk {ro['jerk_cost']:.6e} iters {ro['iterations']} evals {ro['evaluations']} st {ro['status']} ({dto*1e3:.1f} ms cpu)"
                  f" rel {abs(r['final_cost'][i]-ro['final_cost'])/abs(ro['final_cost']):.2e}", flush=True)
        print("   argmin", batch.argmin(), int(np.argmin(r['final_cost'])))
    batch.close()

if __name__ == "__main__":
    which = sys.argv[1:] or ["small", "c3", "c4"]
    if "small" in which:
        check(synth.make_batch(1, 4, 8, 12, 16), "small(M8,K16)", solve_tol=1e-5)
        check(synth.config_problems("C1", 1), "C1", solve_tol=1e-6)
        check(synth.config_problems("C2", 1), "C2", solve_tol=1e-6)
    if "c3" in which:
        check(synth.config_problems("C3", 64), "C3", solve_tol=1e-5)
    if "c4" in which:
        check(synth.config_problems("C4", 1024), "C4")
    if "c4solve" in which:
        check(synth.config_problems("C4", 256), "C4/256", solve_tol=1e-5)
