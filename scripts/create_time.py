"""Batch creation time (host geometry vs the batched device extraction): FRX_HOST_GEOMETRY=1 forces the host path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_racing_b200 import api, synth
cfg = sys.argv[1] if len(sys.argv) > 1 else "C5"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
probs = synth.config_problems(cfg, B)
api.Batch(probs[:4]).close()  # context + kernels loaded
for mode in ("1", "0"):
    os.environ["FRX_HOST_GEOMETRY"] = mode
    t = time.perf_counter(); b = api.Batch(probs); dt = time.perf_counter() - t
    print(f"{cfg} B={B} {'host' if mode == '1' else 'device'} geometry: create {dt*1e3:.1f} ms = {dt/B*1e6:.1f} us/trajectory", flush=True)
    b.close()
