"""End-to-end (host pinned buffers) evaluation time, C4."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_racing_b200 import api, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
b = api.Batch(synth.config_problems("C4", B))
x = b.initial_guess() + 0.5*np.random.default_rng(1).normal(size=b.total_x)
c_ref, g_ref = b.eval(x)
xp, cp, gp = b.host_buffers(); xp[:] = x
for _ in range(3): b.eval_pinned()
assert np.array_equal(cp, c_ref) and np.array_equal(gp, g_ref)
t = time.perf_counter(); N = 30
for _ in range(N): b.eval_pinned()
dt = (time.perf_counter() - t) / N
print(f"chunks={os.environ.get('FRX_E2E_CHUNKS','default')}: {dt*1e3:.3f} ms/step -> {B/dt:.3e} evals/s (device-timed {b.last_ms():.3f} ms)")
