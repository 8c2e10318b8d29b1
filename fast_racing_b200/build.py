"""Builds fast_racing_b200/libfrx.so (host C++ + sm_100a CUDA) in-tree with nvcc.

Called by __graft_entry__.build(); the .so is git-ignored but travels to the GPU box with the
snapshot.  No JIT, no torch extension machinery: the product is a plain C-ABI shared library.

The device sources are compiled twice: the fast build (FMA contraction, reciprocal pivots, parallel
reductions; namespace frx) and the reference-order build (-DFRX_EXACT --fmad=false; namespace
frx_exact), which reproduces the reference's operation order bit for bit (DESIGN.md "Numerics").
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# FRX_BUILD_TAG=dbg (with FRX_NVCC_EXTRA=...) builds an instrumented variant next to the product: libfrx_dbg.so
_TAG = os.environ.get("FRX_BUILD_TAG", "")
OBJ = os.path.join(HERE, "build" + ("_" + _TAG if _TAG else ""))
OUT = os.path.join(HERE, "libfrx" + ("_" + _TAG if _TAG else "") + ".so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-O3,-pthread", "-x", "cu"]
# (source, object suffix, extra flags)
UNITS = [
    ("frx_capi.cu", "", []),
    ("eval_kernels.cu", "", []),
    ("solver.cu", "", []),
    ("geometry_kernels.cu", "", ["--fmad=false"]),  # bit-identical to the host enumeration (host_geometry.cpp)
    ("host_model.cpp", "", []),
    ("host_geometry.cpp", "", []),
    ("corridor.cpp", "", []),
    ("corridor_kernels.cu", "", ["--fmad=false"]),  # same point selection as the host path (corridor.cpp)
    ("eval_kernels.cu", "_exact", ["-DFRX_EXACT=1", "--fmad=false"]),
    ("solver.cu", "_exact", ["-DFRX_EXACT=1", "--fmad=false"]),
]


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "frx.h"), __file__]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    newest_hdr = max(os.path.getmtime(d) for d in _deps() if d.endswith((".cuh", ".hpp", ".h", ".py")))
    procs, objs = [], []
    for src, suffix, extra in UNITS:
        obj = os.path.join(OBJ, os.path.splitext(src)[0] + suffix + ".o")
        objs.append(obj)
        srcp = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(srcp), newest_hdr):
            continue
        cmd = [nvcc] + ARCH + COMMON + extra + os.environ.get("FRX_NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else []) + ["-c", srcp, "-o", obj]
        procs.append((src + suffix, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for name, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(f"---- {name} ----\n{out}")
            failed = True
        elif verbose:
            sys.stderr.write(f"---- {name} ----\n{out}")
    if failed:
        raise RuntimeError("nvcc failed building libfrx.so")
    r = subprocess.run([nvcc] + ARCH + ["-shared", "-o", OUT] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
