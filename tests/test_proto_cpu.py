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
