"""Time full batched solves: ms per tick, solves/s."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_racing_b200 import api, synth
cfg = sys.argv[1] if len(sys.argv) > 1 else "C4"
for B in [int(a) for a in sys.argv[2:]] or [256]:
    b = api.Batch(synth.config_problems(cfg, B))
    t = time.time(); r = b.optimize(1e-5); dt = time.time() - t
    print(f"{cfg} B={B}: wall {dt:.2f}s device {r['device_ms']:.0f} ms ticks {r['ticks']} -> {r['device_ms']/r['ticks']*1e3:.0f} us/tick, {B/dt:.1f} solves/s, "
          f"iters mean {r['iterations'].mean():.0f} max {r['iterations'].max()} evals mean {r['evaluations'].mean():.0f} ok {np.mean(r['status']>=0):.3f}", flush=True)
    b.close()
