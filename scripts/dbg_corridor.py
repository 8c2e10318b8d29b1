import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_racing_b200 import api, synth
BBOX = np.array([4.0, 4.0, 2.5])
rng = np.random.default_rng(5)
centres = synth.corridor_centres(rng, 24)
obs = synth.obstacle_cloud(5, centres, 60000)
segs = np.array([np.r_[a, b] for a, b in zip(centres[:-1], centres[1:])])
cloud = api.Cloud(obs)
faces, cnt, ell = cloud.dilate_batch(segs, BBOX)
for s, seg in enumerate(segs):
    f, Cm, d = api.corridor_dilate(obs, seg[:3], seg[3:], BBOX)
    same = cnt[s] == len(f) and np.array_equal(faces[s][:-6, 3:], f[:-6, 3:])
    de = np.abs(ell[s] - np.r_[Cm.reshape(-1), d]).max()
    if not same or de > 1e-12:
        k = 0
        while k < min(cnt[s], len(f)) - 6 and np.array_equal(faces[s][k, 3:], f[k, 3:]): k += 1
        print(f"seg {s}: device {cnt[s]} host {len(f)} faces; first differing face {k}; ellipsoid max diff {de:.3e}")
        print("  host ell", np.r_[Cm.reshape(-1), d]); print("  dev  ell", ell[s])
        if k < min(cnt[s], len(f)): print("  host face", f[k], "\n  dev  face", faces[s][k])
    else:
        print(f"seg {s}: ok ({cnt[s]} faces), ell diff {de:.1e}, normal diff {np.abs(faces[s]-f).max():.1e}")
