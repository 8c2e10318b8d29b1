"""Accumulated clock64() phase times of lbfgs_update_kernel (CTA 0 = trajectory 0) over one batched solve.
Needs the instrumented build:  FRX_BUILD_TAG=dbg FRX_NVCC_EXTRA=-DFRX_PHASE_CLK ... ; FRX_LIB=.../libfrx_dbg.so"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_racing_b200 import api, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
b = api.Batch(synth.config_problems("C4", B))
out = (C.c_longlong * 16)()
b.lib.frx_debug_update_clocks.argtypes = [C.POINTER(C.c_longlong), C.c_int]
b.lib.frx_debug_update_clocks(out, 1)
r = b.optimize(1e-5)
b.lib.frx_debug_update_clocks(out, 0)
v = np.array(out[:], dtype=np.float64)
names = ["dg dot", "ls step", "xx/gg dots", "conv + history write + ys/yy", "pass 1 (Gram dots)", "loop 1", "gamma + tv", "loop 2", "pass 2 (combine)",
         "save xp/gp + dg0", "trial x / finish"]
it, ev = int(r["iterations"][0]), int(r["evaluations"][0])
print(f"trajectory 0: {it} iterations, {ev} evaluations, ticks {r['ticks']}; total {v[:11].sum()/1e6:.1f} Mcycles")
for n, c in zip(names, v[:11]):
    print(f"  {n:32s} {c/1e6:8.2f} Mcyc  {c/max(it,1):9.0f} cyc/iteration  {100*c/v[:11].sum():5.1f}%")
uk = r["update_kernel"]
print("update kernel sampled mean %.1f us, %.0f GB/s" % (1e3*uk["device_ms"]/uk["sampled_launches"], uk["history_bytes"]/uk["device_ms"]/1e6))
