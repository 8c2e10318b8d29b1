"""Fast mode against reference-order mode (bit-identical to the reference) on every trajectory of C4 at a perturbed point."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_racing_b200 import api, synth
B = 1024
b = api.Batch(synth.config_problems("C4", B))
rng = np.random.default_rng(9)
x = b.initial_guess() + 0.3 * rng.normal(size=b.total_x)
c1, g1 = b.eval(x)
b.set_reference_order(1)
c2, g2 = b.eval(x)
b.set_reference_order(0)
rc = np.abs(c1 - c2) / np.abs(c2)
rg = np.array([np.abs(g1[b.x_off[i]:b.x_off[i+1]] - g2[b.x_off[i]:b.x_off[i+1]]).max() / np.abs(g2[b.x_off[i]:b.x_off[i+1]]).max() for i in range(B)])
print("cost rel diff fast vs reference-order: median %.2e max %.2e (traj %d)" % (np.median(rc), rc.max(), rc.argmax()))
print("grad rel diff: median %.2e max %.2e (traj %d)" % (np.median(rg), rg.max(), rg.argmax()))
print("worst 5 grad:", np.sort(rg)[-5:])
