"""Batched solves with queue-fed solver slots: P problems, S slots in flight.  usage: slot_sweep.py CFG P S1 S2 ..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_racing_b200 import api, synth
cfg, P = sys.argv[1], int(sys.argv[2])
b = api.Batch(synth.config_problems(cfg, P))
base = None
for S in [int(a) for a in sys.argv[3:]]:
    b.set_solver_slots(S)
    b.optimize(1e-5)  # warm (graph instantiation, clocks)
    t = time.time(); r = b.optimize(1e-5); dt = time.time() - t
    same = "" if base is None else f" same-as-first {np.array_equal(base['x'], r['x']) and np.array_equal(base['status'], r['status'])}"
    base = base or r
    print(f"{cfg} P={P} slots={S or P}: wall {dt*1e3:.0f} ms device {r['device_ms']:.0f} ms ticks {r['ticks']} -> "
          f"{r['device_ms']/r['ticks']*1e3:.0f} us/tick, {P/dt:.1f} solves/s, evals mean {r['evaluations'].mean():.0f} "
          f"max {r['evaluations'].max()} ok {np.mean(r['status']>=0):.3f}{same} | update kernel "
          f"{1e3*r['update_kernel']['device_ms']/max(r['update_kernel']['sampled_launches'],1):.0f} us "
          f"{r['update_kernel']['history_bytes']/max(r['update_kernel']['device_ms'],1e-9)/1e6:.0f} GB/s", flush=True)
b.close()
