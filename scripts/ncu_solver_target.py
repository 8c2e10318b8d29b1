"""Target for ncu captures of the solver's update kernel: one batched solve (C4, B trajectories), capped in ticks.
usage: FRX_SOLVER_GRAPH=0 FRX_SOLVER_MAX_TICKS=1500 ncu -k regex:lbfgs_update -s 1200 -c 2 ... python scripts/ncu_solver_target.py 256"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_racing_b200 import api, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
b = api.Batch(synth.config_problems("C4", B))
r = b.optimize(1e-5)
print("ticks", r["ticks"], "iterations mean", r["iterations"].mean())
b.close()
