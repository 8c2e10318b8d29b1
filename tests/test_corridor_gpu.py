"""Corridor generation, GPU tier: the batched segment step on the device (one CTA per segment, obstacle cloud
resident in HBM) against the host path — which the CPU tier pins to the reference's own decomposition code — and the whole chain
front-end path -> corridor -> SE3GCOPTER solve on the GPU."""
import numpy as np
import pytest

from fast_racing_b200 import api, synth

pytestmark = pytest.mark.gpu
BBOX = np.array([4.0, 4.0, 2.5])


def test_device_dilate_batch_matches_host_path():
    segs, clouds = [], None
    rng = np.random.default_rng(5)
    centres = synth.corridor_centres(rng, 24)
    obs = synth.obstacle_cloud(5, centres, 60000)
    segs = np.array([np.r_[a, b] for a, b in zip(centres[:-1], centres[1:])])
    cloud = api.Cloud(obs)
    faces, cnt, ell = cloud.dilate_batch(segs, BBOX)
    assert (cnt > 6).all()
    swapped = 0
    for s, seg in enumerate(segs):
        f, Cm, d = api.corridor_dilate(obs, seg[:3], seg[3:], BBOX)
        assert cnt[s] == len(f), (s, cnt[s], len(f))
        assert np.abs(ell[s] - np.r_[Cm.reshape(-1), d]).max() <= 1e-12
        assert np.abs(faces[s][-6:] - f[-6:]).max() <= 1e-12  # the six box planes
        # The obstacle points the tangent planes pass through: the same SET, bit for bit.  The shrunk ellipsoid touches two or
        # three cloud points at distance 1 -+ 1e-16, so which of them is "closest" first depends on the last bit of the
        # rotation matrix (device sin/cos/atan2 vs glibc): such neighbours may swap places, the polytope is the same.
        key = lambda a: np.lexsort((a[:, 5], a[:, 4], a[:, 3]))
        fd, fh = faces[s][:-6], f[:-6]
        swapped += int(not np.array_equal(fd[:, 3:], fh[:, 3:]))
        fd, fh = fd[key(fd)], fh[key(fh)]
        assert np.array_equal(fd[:, 3:], fh[:, 3:]), s
        assert np.abs(fd[:, :3] - fh[:, :3]).max() <= 1e-12
    assert swapped <= len(segs) // 2
    # bbox = 0: no box planes, the whole cloud is considered (DecompBase::add_local_bbox returns early)
    f0, c0, _ = cloud.dilate_batch(segs[:2], np.zeros(3))
    for s in range(2):
        f, _, _ = api.corridor_dilate(obs, segs[s][:3], segs[s][3:], np.zeros(3))
        assert c0[s] == len(f) and np.array_equal(np.sort(f0[s][:, 3:], axis=0), np.sort(f[:, 3:], axis=0))
    cloud.close()


def test_path_to_trajectory_on_the_gpu():
    """MavGlobalPlanner::plan's back half end to end: path + obstacle cloud -> corridor (walk on the host, segment steps as the
    reference runs them) -> frx_batch_create / optimize: a feasible trajectory from the first to the last path point."""
    rng = np.random.default_rng(21)
    centres = synth.corridor_centres(rng, 10)
    obs = synth.obstacle_cloud(21, centres, 30000)
    path = synth.sample_path(centres)
    off, faces = api.corridor_generate(obs, path, BBOX, 3.0)
    assert len(off) - 1 >= 3
    p = api.Problem(np.r_[path[0], np.zeros(6)], np.r_[path[-1], np.zeros(6)], off, faces, dict(api.ZHANGJIAJIE, itg_spaces=16))
    b = api.Batch([p])
    r = b.optimize(1e-5)
    assert np.isfinite(r["final_cost"]).all() and r["jerk_cost"][0] > 0
    # every piece stays inside its polytope at its junctions (positions = constant coefficients, low order last)
    co = r["coeffs"].reshape(-1, 6, 3)
    tab = b.tables(0)
    for i in range(len(co)):
        f = faces[off[tab["idx_hs"][i]]:off[tab["idx_hs"][i] + 1]]
        f = np.c_[f[:, :3] / np.linalg.norm(f[:, :3], axis=1, keepdims=True), f[:, 3:]]
        assert ((co[i, 5] - f[:, 3:]) * f[:, :3]).sum(1).max() < 0.3  # start point of the piece, up to the soft penalty's slack
    b.close()


def test_batched_walks_on_the_device_match_the_host_walk():
    """frx_cloud_generate_batch: the corridors of several candidate paths grown in lock-step (one batched device step per walk
    step) are the polytopes the host walk builds for each path alone (faces as sets per polytope: see the near-tie note above)."""
    rng = np.random.default_rng(31)
    centres = synth.corridor_centres(rng, 14)
    obs = synth.obstacle_cloud(31, centres, 50000)
    base = synth.sample_path(centres)
    paths = [base, base[: len(base) // 2], base[len(base) // 3:], base[::2]]
    cloud = api.Cloud(obs)
    got = cloud.generate_batch(paths, BBOX, 3.0)
    for path, (off, faces) in zip(paths, got):
        off_h, faces_h = api.corridor_generate(obs, path, BBOX, 3.0)
        assert np.array_equal(off, off_h)
        for k in range(len(off) - 1):
            a, b = faces[off[k]:off[k + 1]], faces_h[off[k]:off[k + 1]]
            key = lambda m: np.lexsort((m[:, 5], m[:, 4], m[:, 3], m[:, 2]))
            assert np.abs(a[key(a)] - b[key(b)]).max() <= 1e-12
    cloud.close()
