"""Seeded synthetic safe-flight corridors.

The reference's tracks cannot be reproduced offline (gate poses come from a live AirSim RPC,
se3_node_cpu.cpp:41-78), so the benchmark configurations use stand-in corridors: a chain of convex
polytopes around a smooth random 3-D curve.  Polytope i wraps the segment c_i -> c_{i+1}: six faces of a
box aligned with the segment (floor and ceiling are the same two global planes for every polytope, as in
MinCoPlan_CPU.cpp:85-91) plus (F - 6) random tangent planes that keep both segment end points strictly
inside, so consecutive intersections are never empty.  Output format = the reference's hPolys: one
(nx,ny,nz,px,py,pz) row per face (one column of the 6xF matrix, MinCoPlan_CPU.cpp:101).
"""
import math

import numpy as np

from .api import MODERNCITY, ZHANGJIAJIE, Problem


def corridor_centres(rng, n_seg, step=3.0, step_jitter=0.5, max_turn_deg=30.0, z_range=(0.8, 2.2), start=None):
    c = np.zeros((n_seg + 1, 3))
    c[0] = start if start is not None else np.array([0.0, 0.0, 0.5 * (z_range[0] + z_range[1])])
    heading = rng.uniform(-math.pi, math.pi)
    for i in range(n_seg):
        heading += math.radians(rng.uniform(-max_turn_deg, max_turn_deg))
        L = step + rng.uniform(-step_jitter, step_jitter)
        dz = rng.uniform(-0.3, 0.3)
        nxt = c[i] + np.array([L * math.cos(heading), L * math.sin(heading), dz])
        nxt[2] = min(max(nxt[2], z_range[0]), z_range[1])
        c[i + 1] = nxt
    return c


def segment_polytope(rng, a, b, n_faces, lateral=3.0, along_margin=1.5, floor=0.0, ceiling=3.0, tube=None):
    """Faces (n_faces x 6) of the polytope around segment a->b."""
    mid = 0.5 * (a + b)
    d = b - a
    L = np.linalg.norm(d)
    t = d / L
    th = np.array([t[0], t[1], 0.0])
    th /= np.linalg.norm(th)
    lat = np.array([-th[1], th[0], 0.0])
    faces = []
    half = 0.5 * L + along_margin
    w = lateral if tube is None else tube
    faces.append(np.r_[th, mid + half * th])
    faces.append(np.r_[-th, mid - half * th])
    faces.append(np.r_[lat, mid + w * lat])
    faces.append(np.r_[-lat, mid - w * lat])
    faces.append(np.r_[0.0, 0.0, 1.0, mid[0], mid[1], ceiling])
    faces.append(np.r_[0.0, 0.0, -1.0, mid[0], mid[1], floor])
    r_end = 0.5 * L
    while len(faces) < n_faces:
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        # keep both end points at least 0.6 m inside: plane distance from mid >= |n.t| * L/2 + 0.6
        dist = abs(float(n @ t)) * r_end + rng.uniform(0.8, 2.2)
        faces.append(np.r_[n, mid + dist * n])
    return np.array(faces)


def make_problem(seed, n_poly, n_faces=12, itg_spaces=32, params=None, gates=(), chi=None, tube=1.2, start=None,
                 centres=None):
    """One candidate trajectory: n_poly polytopes (=> n_poly pieces, gridRes = inf), rest-to-rest."""
    rng = np.random.default_rng(seed)
    par = dict(ZHANGJIAJIE if params is None else params)
    par["itg_spaces"] = int(itg_spaces)
    if chi is not None:
        par["chi"] = tuple(chi)
    c = corridor_centres(rng, n_poly, start=start) if centres is None else np.asarray(centres, dtype=float)
    faces, off = [], [0]
    for i in range(n_poly):
        nf = n_faces if np.isscalar(n_faces) else int(n_faces[i])
        f = segment_polytope(rng, c[i], c[i + 1], nf, tube=(tube if i in gates else None))
        faces.append(f)
        off.append(off[-1] + len(f))
    ini = np.zeros((3, 3))
    fin = np.zeros((3, 3))
    ini[:, 0] = c[0]
    fin[:, 0] = c[-1]
    # column-major 3x3 [p v a]
    return Problem(ini.T.reshape(-1), fin.T.reshape(-1), np.array(off, np.int32), np.concatenate(faces, 0), par)


def make_batch(seed0, count, n_poly, n_faces=12, itg_spaces=32, params=None, **kw):
    return [make_problem(seed0 + i, n_poly, n_faces, itg_spaces, params, **kw) for i in range(count)]


def perturb_track(base_centres, seed, sigma=0.25):
    """C5-style Monte-Carlo perturbation of a base track's gate centres (end points fixed)."""
    rng = np.random.default_rng(seed)
    c = np.array(base_centres, dtype=float)
    c[1:-1] += rng.normal(scale=sigma, size=c[1:-1].shape)
    return c


def obstacle_cloud(seed, centres, n_points=20000, clearance=0.7, margin=5.0, z_range=(0.0, 3.0)):
    """A random obstacle cloud around a path that leaves a tube of radius `clearance` around the poly-line free: the kind of
    input MavGlobalPlanner::plan hands to EllipsoidDecomp3D::set_obs (MinCoPlan_CPU.cpp:38), for the corridor-generation tests."""
    rng = np.random.default_rng(seed)
    c = np.asarray(centres, dtype=float)
    lo = np.r_[c[:, :2].min(0) - margin, z_range[0]]
    hi = np.r_[c[:, :2].max(0) + margin, z_range[1]]
    pts = rng.uniform(lo, hi, size=(n_points, 3))
    dmin = np.full(len(pts), np.inf)
    for a, b in zip(c[:-1], c[1:]):
        d = b - a
        t = np.clip(((pts - a) @ d) / (d @ d), 0.0, 1.0)
        dmin = np.minimum(dmin, np.linalg.norm(pts - (a + t[:, None] * d), axis=1))
    return np.ascontiguousarray(pts[dmin > clearance])


def sample_path(centres, spacing=0.25):
    """Densely sampled poly-line through the centres (what the front end's getSamplePath returns, MinCoPlan_CPU.cpp:19-35)."""
    c = np.asarray(centres, dtype=float)
    out = [c[0]]
    for a, b in zip(c[:-1], c[1:]):
        n = max(1, int(np.ceil(np.linalg.norm(b - a) / spacing)))
        for k in range(1, n + 1):
            out.append(a + (b - a) * k / n)
    return np.array(out)


CONFIGS = {
    # BASELINE.json configs[0..4]; B is the full-size batch, tests use fewer.
    "C1": dict(n_poly=8, n_faces=10, itg_spaces=48, params=ZHANGJIAJIE, chi=(0.0, 1.0e4, 1.0e4, 1.0e4), B=1),
    "C2": dict(n_poly=24, n_faces=(8, 30), itg_spaces=48, params=ZHANGJIAJIE, B=1),
    "C3": dict(n_poly=32, n_faces=12, itg_spaces=16, params=ZHANGJIAJIE, B=64),
    "C4": dict(n_poly=64, n_faces=12, itg_spaces=32, params=ZHANGJIAJIE, B=1024),
    "C5": dict(n_poly=24, n_faces=12, itg_spaces=48, params=MODERNCITY, B=8192),
}


def config_problems(name, count=None, seed0=1, itg_spaces=None):
    cfg = CONFIGS[name]
    B = cfg["B"] if count is None else count
    K = cfg["itg_spaces"] if itg_spaces is None else itg_spaces
    out = []
    if name == "C5":
        base = corridor_centres(np.random.default_rng(12345), cfg["n_poly"])
        for i in range(B):
            out.append(make_problem(seed0 + i, cfg["n_poly"], cfg["n_faces"], K, cfg["params"],
                                    centres=perturb_track(base, seed0 + i)))
        return out
    for i in range(B):
        nf = cfg["n_faces"]
        if not np.isscalar(nf):
            r = np.random.default_rng(1000 + seed0 + i)
            nf = r.integers(nf[0], nf[1] + 1, size=cfg["n_poly"])
        gates = range(1, cfg["n_poly"], 2) if name in ("C1",) else (range(2, cfg["n_poly"], 3) if name == "C2" else ())
        out.append(make_problem(seed0 + i, cfg["n_poly"], nf, K, cfg["params"], gates=tuple(gates), chi=cfg.get("chi")))
    return out
