"""Slot algebra of the experimental NP-pivots-per-trip elimination (csrc/eval_kernels.cu, FRX_K1_NP; not built by default):
the numpy emulation must reproduce a dense no-pivot LU + forward substitution for NP = 2 (today's kernel), 3 and 6."""
import importlib.util
import os
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_np_pivot_trip_emulation_matches_dense_lu():
    spec = importlib.util.spec_from_file_location("k1_emu", os.path.join(ROOT, "scripts", "proto", "k1_piece_trip_emulation.py"))
    emu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(emu)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # 1/0 for the "next pivot" of the zero rows after the matrix (never used)
        for NP in (2, 3, 6):
            for M, seed in ((1, 0), (3, 1), (9, 2)):
                assert emu.check(M, seed, NP) < 1e-10


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "scripts", "proto", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_cyclic_reduction_emulation_matches_dense_solve():
    """pcr_solve (csrc/eval_kernels.cu): the numpy emulation of its slot algebra — symmetric storage, lower coupling only,
    A_1 = 0 — against a dense solve of the reduced MINCO system, for every node count the kernel takes (1..64) including the
    non-powers of two, with durations spread over two decades."""
    import numpy as np
    pcr = _load("pcr_reduced")
    rng = np.random.default_rng(3)
    for M in (2, 3, 4, 5, 6, 7, 9, 16, 17, 31, 33, 48, 64, 65):
        T = np.exp(rng.normal(0, 1.0, size=M))
        N = M - 1
        D, A = pcr.build(T)
        R = rng.normal(size=(N + 2, 6)) * 100
        Y = pcr.pcr(D, A, R, N)
        S = pcr.dense(D, A, N)
        for d in range(3):
            rhs = np.zeros(2 * N)
            rhs[0::2], rhs[1::2] = R[1:N + 1, d], R[1:N + 1, 3 + d]
            y = np.linalg.solve(S, rhs)
            got = np.zeros(2 * N)
            got[0::2], got[1::2] = Y[1:N + 1, d], Y[1:N + 1, 3 + d]
            assert np.abs(got - y).max() <= 1e-11 * np.abs(y).max(), (M, d)
