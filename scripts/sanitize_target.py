"""Small workload for compute-sanitizer: ragged batch, one evaluation + a short solve (fast build: cyclic reduction in K1/K3, TMA-streamed
direction update once the history has pairs), the same solve through 3 solver slots, one evaluation in reference-order mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_racing_b200 import api, synth
probs = synth.make_batch(6, 5, 9, 10, 8) + synth.config_problems("C3", 2)
batch = api.Batch(probs)
x = batch.initial_guess() + 0.3 * np.random.default_rng(0).normal(size=batch.total_x)
c, g = batch.eval(x)
print("eval ok", np.isfinite(c).all(), np.isfinite(g).all())
r = batch.optimize(float(sys.argv[1]) if len(sys.argv) > 1 else 1e-2)
print("solve status", r["status"], "iters", r["iterations"])
batch.set_solver_slots(3)  # queue-fed slots: 7 trajectories through 3 slots
r2 = batch.optimize(float(sys.argv[1]) if len(sys.argv) > 1 else 1e-2)
print("slot-mode solve identical", bool(np.array_equal(r["x"], r2["x"])))
batch.set_solver_slots(0)
batch.set_reference_order(True)
c2, g2 = batch.eval(x)
print("reference-order eval rel diff", float(np.abs(c2 / c - 1).max()))
