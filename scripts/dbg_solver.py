import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from fast_racing_b200 import api, synth
import oracle.binding as ob
probs = synth.make_batch(50, 6, 6, 10, 8)
batch = api.Batch(probs)
r = batch.optimize(1e-13)
print("status", r["status"], "iters", r["iterations"], "evals", r["evaluations"])
x = r["x"]
c, g = batch.eval(x)
for i, p in enumerate(probs):
    off, verts = batch.vpolys(i)
    o = ob.Oracle(p.with_vpolys(off, verts))
    p0, p1 = batch.x_off[i], batch.x_off[i + 1]
    f, go = o.eval(x[p0:p1])
    ge = g[p0:p1]
    print(i, "cost rel", abs(f - c[i]) / abs(f), "grad rel", np.linalg.norm(ge - go) / np.linalg.norm(go), "|g|", np.linalg.norm(go), "argmax", int(np.argmax(np.abs(ge - go))), "n", p1 - p0)
