"""Drive the reference L-BFGS with the CUDA evaluator for ONE trajectory and compare every evaluation with the oracle."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_racing_b200 import api, synth
import oracle.binding as ob
idx = int(sys.argv[1]) if len(sys.argv) > 1 else 3
probs = synth.make_batch(50, 6, 6, 10, 8)
full = api.Batch(probs)
off, verts = full.vpolys(idx)
p = probs[idx].with_vpolys(off, verts)
batch = api.Batch([p]); o = ob.Oracle(p)
R = ob.ref_lbfgs(); assert R is not None
x = batch.initial_guess().copy(); n = len(x)
worst = [0.0, 0.0, 0]; cnt = [0]
def cb(inst, xx, g, nn):
    xv = np.ctypeslib.as_array(xx, shape=(n,)).copy()
    c, gg = batch.eval(xv)
    fo, go = o.eval(xv)
    ec = abs(c[0] - fo) / max(abs(fo), 1e-300); eg = np.linalg.norm(gg - go) / max(np.linalg.norm(go), 1e-300)
    cnt[0] += 1
    if ec > 1e-10 or eg > 1e-8:
        if worst[2] < 5:
            print(f"eval {cnt[0]}: cost rel {ec:.3e} grad rel {eg:.3e} f={fo:.6e} argmax {int(np.argmax(np.abs(gg-go)))} n={n} diff {np.abs(gg-go).max():.3e}")
            np.save(f"/root/repo/gpurun_out/badx_{idx}_{worst[2]}.npy", xv)
        worst[2] += 1
    worst[0] = max(worst[0], ec); worst[1] = max(worst[1], eg)
    np.ctypeslib.as_array(g, shape=(n,))[:] = gg
    return float(c[0])
f = C.c_double()
CB = ob.EVAL_T(cb)
rc = R.ref_lbfgs_optimize(n, x.ctypes.data_as(C.POINTER(C.c_double)), C.byref(f), CB, None, 128, 3, 1e-13, 1e-16, 1e-32, 0)
print("rc", rc, "evals", cnt[0], "worst cost", worst[0], "worst grad", worst[1], "bad", worst[2], "f", f.value)
