"""CPU tier: the product's host side (no GPU): exported C ABI, geometry, integer tables, starting point,
error behaviour, and the world_size-2 gloo path of the arg-min gather."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

from fast_racing_b200 import api, synth
from helpers import with_product_vpolys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_of_the_header():
    hdr = open(os.path.join(ROOT, "include", "frx.h")).read()
    declared = sorted(set(re.findall(r"\b(frx_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 18
    lib = C.CDLL(api.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"libfrx.so does not export {name}"
    assert set(api.EXPORTED_SYMBOLS) <= set(declared)
    api.load_library()
    assert b"sm_100a" in api.load_library().frx_version()


@pytest.mark.parametrize("seed", range(6))
def test_vertex_enumeration_matches_brute_force(oracle_mod, seed):
    """Every polytope AND every consecutive intersection (which contains exactly duplicated floor/ceiling planes)."""
    p = synth.make_problem(seed, 5, np.array([6, 9, 12, 20, 30]), 8, gates=(1,))
    F = p.faces.reshape(-1, 6)
    fo = p.face_off
    for i in range(p.n_poly):
        cases = [F[fo[i]:fo[i + 1]]]
        if i + 1 < p.n_poly:
            cases.append(F[fo[i]:fo[i + 2]])
        for f in cases:
            v = api.enumerate_vertices(f)
            fn = f.copy()
            fn[:, :3] /= np.linalg.norm(fn[:, :3], axis=1, keepdims=True)
            bf = oracle_mod.brute_force_vertices(fn)
            assert len(v) == len(bf)
            d = np.abs(v[:, None, :] - bf[None, :, :]).max(-1)
            assert d.min(1).max() < 1e-9 and d.min(0).max() < 1e-9
            # all vertices satisfy all half-spaces
            assert ((v[:, None, :] - fn[None, :, 3:]) * fn[None, :, :3]).sum(-1).max() < 1e-9


def test_infeasible_and_unbounded_geometry_is_reported():
    p = synth.make_problem(1, 3, 8, 8)
    faces = p.faces.reshape(-1, 6).copy()
    # move polytope 1 far away: intersection 0^1 is empty -> SE3GCOPTER::setup returns false (CPU.hpp:1118-1121)
    faces[p.face_off[1]:p.face_off[2], 3:] += 1000.0
    bad = api.Problem(p.ini, p.fin, p.face_off, faces, p.params)
    with pytest.raises(api.FrxError) as e:
        api.host_model(bad)
    assert e.value.status == api.FRX_ERR_INFEASIBLE_GEOMETRY
    # an unbounded polytope (drop the ceiling) has no V-representation either
    f0 = p.faces.reshape(-1, 6)[p.face_off[0]:p.face_off[1]]
    open_top = np.array([r for r in f0 if not (abs(r[2]) > 0.99 and r[2] > 0)])
    with pytest.raises(api.FrxError):
        api.enumerate_vertices(open_top[:5])


@pytest.mark.parametrize("grid_res", [math.inf, 2.0, 0.7])
def test_index_tables_bit_exact_vs_oracle(oracle_mod, grid_res):
    p = synth.make_problem(4, 6, 12, 8)
    p = api.Problem(p.ini, p.fin, p.face_off, p.faces, dict(p.params, grid_res=grid_res))
    hm = api.host_model(p)
    o = oracle_mod.Oracle(with_product_vpolys(p))
    tab = o.tables()
    assert (hm["n"], hm["pieces"], hm["dim_t"], hm["dim_p"]) == (o.n, o.pieces, o.dim_t, o.dim_p)
    for k in ("intervals", "idx_hs", "idx_vs", "cons"):
        assert np.array_equal(hm[k], tab[k]), k
    if grid_res == math.inf:
        assert (hm["intervals"] == 1).all() and hm["pieces"] == p.n_poly
    else:
        assert hm["pieces"] > p.n_poly
    assert np.array_equal(hm["x0"], o.initial_guess())  # setInitial + backwardT + backwardP


def test_hard_time_and_exp_diffeo_tables(oracle_mod):
    p = synth.make_problem(5, 5, 10, 8)
    for rho, tt, c2 in ((0.0, 6.0, 1), (0.0, 6.0, 0), (500.0, 0.0, 0)):
        q = api.Problem(p.ini, p.fin, p.face_off, p.faces, dict(p.params, rho=rho, total_t=tt, c2_diffeo=c2))
        hm = api.host_model(q)
        o = oracle_mod.Oracle(with_product_vpolys(q))
        assert hm["dim_t"] == o.dim_t == (p.n_poly if rho > 0 else p.n_poly - 1)
        assert np.array_equal(hm["x0"], o.initial_guess())


def test_invalid_arguments():
    lib = api.load_library()
    assert lib.frx_batch_create(None, 0, -1, None) == api.FRX_ERR_INVALID_ARG
    assert lib.frx_batch_count(None) == api.FRX_ERR_INVALID_ARG
    p = synth.make_problem(1, 2, 8, 8)
    bad = api.Problem(p.ini, p.fin, p.face_off, p.faces, dict(p.params, itg_spaces=0))
    with pytest.raises(api.FrxError):
        api.host_model(bad)


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.FrxError) as e:
        api.Batch([synth.make_problem(1, 2, 8, 8)])
    assert e.value.status == api.FRX_ERR_CUDA


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    from fast_racing_b200.dist import global_argmin, shard_range

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shard_range(11, rank, world)
    costs = np.array([5.0, 3.0, 9.0, 3.0, 7.0, 1.5, 8.0, 1.5, 6.0, 4.0, 2.0])
    local = costs[lo:hi]
    li = int(np.argmin(local))
    c, gi = global_argmin(float(local[li]), lo + li, device=torch.device("cpu"))
    q.put((rank, lo, hi, c, gi))
    dist.destroy_process_group()


def test_argmin_gather_world_size_2_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    res.sort()
    assert (res[0][1], res[0][2], res[1][1], res[1][2]) == (0, 6, 6, 11)  # block partition
    for r in res:
        assert r[3] == 1.5 and r[4] == 5  # ties -> lowest global index, same answer on every rank


def test_update_kernel_ring_geometry_invariants():
    """The streamed L-BFGS direction update (csrc/solver.cu) copies the history with 1-D bulk copies, which need 16-byte aligned
    addresses and sizes and must fit their stage.  The ring configuration (host) and the per-trajectory chunk geometry (the SAME
    __host__ __device__ function the kernel calls) are checked here for every largest-trajectory size the streamed path takes
    (1..2048 variables), small and large batches, and ragged members of a batch: sizes, alignment, coverage, shared-memory budget."""
    lib = api.load_library()
    lib.frx_debug_update_ring.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.frx_debug_update_ring.restype = C.c_int
    out = (C.c_int * 8)()
    M_HIST, STATIC_SMEM, SM_SMEM, RESERVED = 128, 14592, 233472, 1024  # history depth; update kernel's static smem; per-SM; per-CTA
    streamed_sizes = 0
    for batch in (1, 296, 297, 4096):
        for max_n in list(range(1, 300)) + list(range(300, 2049, 7)) + [1183, 1184, 2047, 2048]:
            members = sorted(v for v in {max_n, 1, 2, 3, max(1, max_n // 2), max(1, max_n - 1)} if v <= max_n)
            for n in members:
                assert lib.frx_debug_update_ring(max_n, n, batch, out) == 0
                nst, ch, smem, W, CT, R2, streamed, fixed = list(out)
                assert streamed == 1, (batch, max_n, n)  # every size up to 2048 takes the streamed path
                streamed_sizes += 1
                ld = (n + 1) & ~1
                assert 2 <= nst <= 8 and ch % 2 == 0 and smem == 8 * (nst * ch + fixed)
                assert smem + STATIC_SMEM <= 227 * 1024  # opt-in limit per CTA
                if batch > 296:  # many trajectories: two CTAs per SM must fit whenever the small configuration was taken
                    if nst == 3:
                        assert 2 * (smem + STATIC_SMEM + RESERVED) <= SM_SMEM, (max_n, smem)
                # pass 1: tiles of 8 rows x W columns, row stride W + 4 doubles in the stage
                assert W % 2 == 0 and (W == ld or W % 32 == 0) and 8 * (W + 4) <= ch
                assert (CT - 1) * W < ld <= CT * W
                for ct in range(CT):
                    wt = min(W, ld - ct * W)
                    assert wt > 0 and wt % 2 == 0 and (ct * W) % 2 == 0  # 16-byte aligned source, 16-byte multiple
                # pass 2: R2 whole rows per chunk
                assert 1 <= R2 <= M_HIST and R2 * ld <= ch
                # the three vectors + partial tiles behind the ring
                assert fixed >= 2 * 8 * 8 * 3 + 3 * ld
    assert streamed_sizes > 5000
    # beyond 8 columns per thread the plain-load path takes over
    assert lib.frx_debug_update_ring(2049, 2049, 1024, out) == 0 and out[2] == 0 and out[6] == 0
