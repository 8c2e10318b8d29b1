"""Batched corridor step on the device vs the host path: B segments against one obstacle cloud."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_racing_b200 import api, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
rng = np.random.default_rng(3)
centres = synth.corridor_centres(rng, B)
obs = synth.obstacle_cloud(3, centres, N, margin=8.0)
segs = np.array([np.r_[a, b] for a, b in zip(centres[:-1], centres[1:])])
bbox = np.array([4.0, 4.0, 2.5])
cloud = api.Cloud(obs)
cloud.dilate_batch(segs[:2], bbox)
t = time.perf_counter(); faces, cnt, _ = cloud.dilate_batch(segs, bbox); dt = time.perf_counter() - t
t = time.perf_counter()
for s in segs[:8]: api.corridor_dilate(obs, s[:3], s[3:], bbox)
dh = (time.perf_counter() - t) / 8
print(f"cloud {len(obs)} points, {len(segs)} segments: device batch {dt*1e3:.2f} ms ({dt/len(segs)*1e6:.0f} us/segment incl. transfers), host {dh*1e3:.2f} ms/segment (1 thread); faces/segment mean {cnt.mean():.1f}")
