"""Short target for ncu: a few hundred ticks of a batched solve (C3-like, B trajectories)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_racing_b200 import api, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = sys.argv[2] if len(sys.argv) > 2 else "C3"
b = api.Batch(synth.config_problems(cfg, B))
r = b.optimize(1e-5)
print("ticks", r["ticks"], "device_ms", r["device_ms"], "iters", r["iterations"].mean())
