"""Per-CTA wall spans of K1 / K3 (instrumented build: FRX_BUILD_TAG=dbg FRX_NVCC_EXTRA=-DFRX_PHASE_CLK; FRX_LIB=.../libfrx_dbg.so):
launch skew, durations, CTAs per SM.  usage: cta_spans.py [B]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fast_racing_b200 import api, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
b = api.Batch(synth.config_problems("C4", B))
x = b.initial_guess() + 0.5 * np.random.default_rng(1).normal(size=b.total_x)
xd = torch.from_numpy(x).cuda(); gd = torch.zeros_like(xd); cd = torch.zeros(B, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream()
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
for k in range(4):
    flush.fill_(k)
    b.eval_device(xd.data_ptr(), cd.data_ptr(), gd.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
out = (C.c_ulonglong * (2 * 3 * 4096))()
b.lib.frx_debug_cta_spans.argtypes = [C.POINTER(C.c_ulonglong)]
assert b.lib.frx_debug_cta_spans(out) == 0
a = np.array(out[:], dtype=np.uint64).reshape(2, 3, 4096)[:, :, :B].astype(np.int64)
for k, name in enumerate(("K1 minco_forward", "K3 minco_backward")):
    s, e, sm = a[k, 0], a[k, 1], a[k, 2]
    t0 = s.min()
    dur = (e - s) / 1e3
    print(f"{name}: grid {B}; first start 0, last start {(s.max()-t0)/1e3:.1f} us, last end {(e.max()-t0)/1e3:.1f} us; CTA duration us: "
          f"min {dur.min():.1f} mean {dur.mean():.1f} p90 {np.percentile(dur,90):.1f} max {dur.max():.1f}; CTAs per SM: max {np.bincount(sm).max()} on {len(np.unique(sm))} SMs")
    order = np.argsort(s)
    q = [0, B // 4, B // 2, 3 * B // 4, B - 1]
    print("   start-ordered CTAs (start us, duration us):", ", ".join(f"({(s[order[i]]-t0)/1e3:.1f}, {dur[order[i]]:.1f})" for i in q))
    late = s - t0 > 5000
    print(f"   CTAs starting later than 5 us after the first: {late.sum()}")
