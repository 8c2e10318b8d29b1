"""Per-phase clock64() differences of one CTA of K1 (minco_forward) and K3 (minco_backward).
Needs a debug build:  FRX_NVCC_EXTRA=-DFRX_PHASE_CLK python -c 'from fast_racing_b200 import build; build.build(force=True)'"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fast_racing_b200 import api, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
b = api.Batch(synth.config_problems("C4", B))
x = b.initial_guess() + 0.5*np.random.default_rng(1).normal(size=b.total_x)
xd = torch.from_numpy(x).cuda(); gd = torch.zeros_like(xd); cd = torch.zeros(B, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream()
for _ in range(5): b.eval_device(xd.data_ptr(), cd.data_ptr(), gd.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
out = (C.c_longlong * 96)()
b.lib.frx_debug_phase_clocks.argtypes = [C.POINTER(C.c_longlong)]
b.lib.frx_debug_phase_clocks(out)
v = np.array(out[:], dtype=np.int64)
names = {0: ["start", "forwardT/P", "node blocks", "block LDL^T + fwd subst", "back subst (+sync)", "Hermite -> c"],
         2: ["start", "bulk staging (TMA)", "Hermite adjoint + rhs", "-", "substitution sweeps / cost", "dJ/dT, dJ/dq", "merge+layerT", "layerP"]}
for k in (0, 2):
    for w in (0, 1):
        m = v[k*32 + w*16: k*32 + w*16 + len(names[k])].astype(np.int64)
        idx = [i for i in range(len(m)) if m[i] != 0]
        parts = [f"{names[k][j]} {m[j] - m[i]}" for i, j in zip(idx[:-1], idx[1:])]
        print(f"K{k+1} warp {w}: total {m[idx[-1]] - m[idx[0]] if idx else 0} cyc | " + " | ".join(parts))
