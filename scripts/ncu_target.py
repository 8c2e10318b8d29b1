"""Short target for ncu: builds the C4 workload (B trajectories) and runs a few batched evaluations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fast_racing_b200 import api, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
K = int(sys.argv[3]) if len(sys.argv) > 3 else 32
batch = api.Batch(synth.config_problems("C4", B, itg_spaces=K))
x = batch.initial_guess() + 0.5 * np.random.default_rng(1234).normal(size=batch.total_x)
xd = torch.from_numpy(x).cuda(); gd = torch.zeros_like(xd); cd = torch.zeros(B, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(reps):
    batch.eval_device(xd.data_ptr(), cd.data_ptr(), gd.data_ptr(), st)
torch.cuda.synchronize()
if "--solve" in sys.argv:
    r = api.Batch(synth.config_problems("C3", 64)).optimize(1e-5)
print("done", float(cd.sum()))
