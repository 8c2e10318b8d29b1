"""Corridor generation, CPU tier: the product's host restatement of LineSegment3D::dilate and of the walk in
MavGlobalPlanner::plan against the reference's OWN decomposition code — decomp_util/{ellipsoid_decomp,line_segment,decomp_base}.h
and decomp_geometry/* compiled verbatim over the Eigen-API shim (oracle/_ref/libref_decomp.so)."""
import ctypes as C
import os

import numpy as np
import pytest

from fast_racing_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libref_decomp.so")
BBOX = np.array([4.0, 4.0, 2.5])  # PolyhedronBox of the reference's parameter files
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))


def _ref():
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libref_decomp.so not built (no /root/reference on this box)")
    return C.CDLL(REF)


def ref_dilate(R, obs, p1, p2, bbox=BBOX):
    faces, ell = np.zeros(6 * 512), np.zeros(12)
    n = R.ref_decomp_dilate(dp(obs), len(obs), dp(np.ascontiguousarray(p1)), dp(np.ascontiguousarray(p2)), dp(bbox), dp(faces), 512, dp(ell))
    return faces[: 6 * n].reshape(n, 6).copy(), ell


def _scene(seed, n_seg=6, n_points=20000):
    rng = np.random.default_rng(seed)
    centres = synth.corridor_centres(rng, n_seg)
    return centres, synth.obstacle_cloud(seed, centres, n_points)


@pytest.mark.parametrize("seed", range(5))
def test_dilate_matches_the_reference_decomposition(seed):
    """Same obstacle points chosen in the same order (the tangent plane of face k passes through the SAME cloud point, bit for
    bit), same number of faces, normals / box planes / ellipsoid equal to rounding."""
    R = _ref()
    centres, obs = _scene(seed)
    for a, b in zip(centres[:-1], centres[1:]):
        f_ref, e_ref = ref_dilate(R, obs, a, b)
        f, Cm, d = api.corridor_dilate(obs, a, b, BBOX)
        assert len(f) == len(f_ref) and len(f) > 6
        assert np.array_equal(f[:-6, 3:], f_ref[:-6, 3:])  # the obstacle points: indexing is bit-exact
        assert np.abs(f[:, :3] - f_ref[:, :3]).max() <= 1e-12 and np.abs(f[-6:, 3:] - f_ref[-6:, 3:]).max() <= 1e-12
        assert np.abs(np.r_[Cm.reshape(-1), d] - e_ref).max() <= 1e-12
        # every tangent plane separates its cloud point's side from the segment: both end points are inside the polytope
        for pt in (a, b):
            assert ((pt - f[:, 3:]) * f[:, :3]).sum(1).max() <= 1e-9


def test_dilate_without_obstacles_and_without_box():
    R = _ref()
    p1, p2 = np.array([0.0, 0.0, 1.0]), np.array([2.0, 1.0, 1.5])
    far = np.array([[100.0, 100.0, 100.0]])
    f_ref, _ = ref_dilate(R, far, p1, p2)
    f, _, _ = api.corridor_dilate(far, p1, p2, BBOX)
    assert len(f) == len(f_ref) == 6 and np.abs(f - f_ref).max() <= 1e-12  # only the local bounding box
    f0, _, _ = api.corridor_dilate(far, p1, p2, np.zeros(3))  # local_bbox_ = 0: no box planes, every point is considered
    f0_ref, _ = ref_dilate(R, far, p1, p2, np.zeros(3))
    assert len(f0) == len(f0_ref) == 1 and np.array_equal(f0[:, 3:], f0_ref[:, 3:])


def _reference_walk(R, obs, path, bbox, map_height, blocked):
    """MinCoPlan_CPU.cpp:44-91 restated around the VERBATIM dilate / Polyhedron3D::inside."""
    polys, i, n = [], 0, len(path)
    while i < n - 1:
        k = i + 1
        while k < n:
            if blocked(path[i], path[k]) or np.linalg.norm(path[i] - path[k]) >= 4.0:
                k -= 1
                break
            k += 1
        if k < i + 1:
            k = i + 1
        if k >= n:
            k = n - 1
        f, _ = ref_dilate(R, obs, path[i], path[k], bbox)
        j = k
        while j < n and R.ref_polyhedron_inside(dp(np.ascontiguousarray(f.reshape(-1))), len(f), dp(np.ascontiguousarray(path[j]))):
            j += 1
        j -= 1
        polys.append(np.vstack([f, [0, 0, 1, 0, 0, map_height], [0, 0, -1, 0, 0, 0]]))
        if j >= n - 1:
            break
        i = int(round((1 * i + 4 * j) // 5))
    return polys


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_corridor_walk_matches_the_reference_walk(seed):
    R = _ref()
    centres, obs = _scene(seed, n_seg=8)
    path = synth.sample_path(centres)
    wall = centres[3]  # "blocked" when the straight line passes within 0.3 m of this centre and is longer than 1.5 m: a stand-in for map_util->isBlocked

    def blocked(a, b):
        d = b - a
        t = np.clip(((wall - a) @ d) / max(d @ d, 1e-300), 0, 1)
        return np.linalg.norm(b - a) > 1.5 and np.linalg.norm(wall - (a + t * d)) < 0.3

    ref_polys = _reference_walk(R, obs, path, BBOX, 3.0, blocked)
    off, faces = api.corridor_generate(obs, path, BBOX, 3.0, blocked)
    assert len(off) - 1 == len(ref_polys) >= 3
    for k, rp in enumerate(ref_polys):
        f = faces[off[k]:off[k + 1]]
        assert len(f) == len(rp) and np.array_equal(f[:-8, 3:], rp[:-8, 3:]) and np.abs(f - rp).max() <= 1e-12
    # the polytopes are a valid corridor for the optimiser: consecutive ones overlap (SE3GCOPTER::setup accepts them)
    p = api.Problem(np.r_[path[0], np.zeros(6)], np.r_[path[-1], np.zeros(6)], off, faces, dict(api.ZHANGJIAJIE, itg_spaces=8))
    assert api.host_model(p)["pieces"] == len(ref_polys)
