import sys; sys.path.insert(0,".")
import numpy as np, torch, ctypes as C
from fast_racing_b200 import api, synth
b = api.Batch(synth.config_problems("C4", 256))
x = b.initial_guess() + 0.5*np.random.default_rng(1).normal(size=b.total_x)
for _ in range(3): b.eval(x)
out = (C.c_longlong*32)(); b.lib.frx_debug_clocks(out)
o = list(out)
for w in (0,1):
    t = o[16*w:16*w+6]; print("warp",w,"sections (cycles):", [t[i+1]-t[i] for i in range(5)], "busy in phases", o[16*w+8])
