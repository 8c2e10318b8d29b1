"""Device-resident batched evaluation time (CUDA events on the caller's stream), C4."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fast_racing_b200 import api, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
b = api.Batch(synth.config_problems("C4", B))
x = b.initial_guess() + 0.5*np.random.default_rng(1).normal(size=b.total_x)
xd = torch.from_numpy(x).cuda(); gd = torch.zeros_like(xd); cd = torch.zeros(B, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream()
flush = torch.empty(64*1024*1024, dtype=torch.float32, device="cuda")
for _ in range(5): b.eval_device(xd.data_ptr(), cd.data_ptr(), gd.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
ts = []
for k in range(20):
    flush.fill_(k); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st); b.eval_device(xd.data_ptr(), cd.data_ptr(), gd.data_ptr(), st.cuda_stream); e1.record(st); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f"B={B} nsub={os.environ.get('FRX_EVAL_SUBBATCHES','default')}: {np.mean(ts)*1e3:.1f} us/eval (min {np.min(ts)*1e3:.1f}) -> {B/np.mean(ts)*1e3:.3e} evals/s; launches {b.eval_launches()}")
ms = b.profile_eval(xd.data_ptr(), cd.data_ptr(), gd.data_ptr(), st.cuda_stream, 20)
print("kernels us: K1 %.1f  K2 %.1f  K3 %.1f" % tuple(1e3 * v for v in ms))
