import glob
import os

import numpy as np

from fast_racing_b200 import api

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    par = {str(k): float(v) for k, v in zip(z["params_keys"], z["params_vals"])}
    par["itg_spaces"] = int(par["itg_spaces"])
    par["c2_diffeo"] = int(par["c2_diffeo"])
    par["chi"] = tuple(z["chi"])
    p = api.Problem(z["ini"], z["fin"], z["face_off"], z["faces"], par, z["vert_off"], z["verts"])
    return p, z


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def with_product_vpolys(problem):
    """Attach the V-polytopes the product's host geometry enumerates (no GPU needed)."""
    off, verts = api.host_vpolys(problem)
    return problem.with_vpolys(off, verts)
