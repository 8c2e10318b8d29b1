#!/usr/bin/env python
"""Benchmark of the hot path: batched L-BFGS cost+gradient evaluations and batched solves (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (libfrx.so)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

A "step" = one batched cost+grad evaluation of every trajectory of the workload (one call of the
lbfgs_evaluate_t boundary per trajectory).  Headline workload = BASELINE.json configs[3] ("C4"): per GPU 1024
candidate trajectories x 64 pieces x 12-face polytopes x 32 quadrature intervals (33 samples) per piece,
evaluated at the optimiser's starting point plus N(0, 0.5^2) noise so that every penalty family (P,V,T,B) is active.
Multi-GPU: the batch shards across ranks (weak scaling: 1024 trajectories per GPU), no data-path
collective; one NCCL all-gather of (cost, index) per solve for the arg-min (timed separately, reported).

Besides the headline the JSON line carries:
  strong     the north_star configuration: 1024 trajectories in TOTAL split over the N ranks
  sweep      K in {8,16,32,64,128}: evals/s, per-kernel times and both roofline fractions of the penalty kernel
  configs    C3 (64 x 32 x 16) evals/s; C1 / C2 single-trajectory solve latency beside the CPU's; C5 (1024 Monte-Carlo
             tracks per GPU, moderncity parameters) solves/s with the cross-rank arg-min
  solves     C4 batched solves/s at the reference's stopping rule, with the CPU's solves/s beside it (cpu_baseline.solves)
  cpu_baseline.cores = threads actually usable (affinity mask and cgroup quota), not os.cpu_count()
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SWEEP_K = (8, 16, 32, 64, 128)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="frx", choices=["frx", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="trajectories per GPU")
    ap.add_argument("--pieces", type=int, default=64)
    ap.add_argument("--faces", type=int, default=12)
    ap.add_argument("--samples", type=int, default=32, help="quadrature intervals per piece (K)")
    ap.add_argument("--noise", type=float, default=0.5, help="sigma of the perturbation of the start point (0.5: all four penalty families active)")
    ap.add_argument("--solve-batch", type=int, default=256, help="trajectories of the C4 solves/sec leg (0 = skip)")
    ap.add_argument("--c5-batch", type=int, default=1024, help="Monte-Carlo tracks per GPU of the C5 solve leg (0 = skip)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline evaluation leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip strong / sweep / configs (headline + solves only)")
    ap.add_argument("--ref-seconds", type=float, default=60.0, help="--impl reference: CPU budget of the whole run (all steps)")
    return ap.parse_args()


def workload_name(a):
    return f"C4: {a.batch} trajectories/GPU x {a.pieces} pieces x {a.faces}-face polytopes x K={a.samples} ({a.samples + 1} samples/piece)"


def algorithmic_bytes(M, F, V, n):
    """SURVEY.md §8(d) per-trajectory byte counts (fp64)."""
    return dict(
        eval_boundary=8 * (2 * n + 1 + 6 * F * M + 3 * V * (M - 1)),          # x -> f, g, constants re-read
        penalty_kernel=8 * (18 * M + M + 6 * F * M + 18 * M + 3 * M),          # c, T, faces in; gdC, gdT, pen, energy out
        pipeline=8 * (2 * (4 * M - 3) + 1 + 6 * F * M) + 8 * 4 * 18 * M,       # (q,T,faces)->(cost,dq,dT) + c/gdC round trip
    )


def algorithmic_flops(M, S, F):
    """M*S*(630 + 40F) + 1700M (all constraints inactive), FMA = 2 flops."""
    return M * S * (630 + 40 * F) + 1700 * M


def effective_cores():
    """Threads this process can actually run: the affinity mask, cut by the cgroup CPU quota (v2 cpu.max / v1 cfs_quota)."""
    aff = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    quota = None
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            quota = float(q) / float(p)
    except Exception:
        try:
            q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = q / p
        except Exception:
            pass
    eff = aff if quota is None else max(1, min(aff, int(math.floor(quota + 1e-9))))
    return {"cpu_count": os.cpu_count(), "affinity": aff, "cgroup_quota": quota, "effective": eff}


class ClockSampler(threading.Thread):
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        sm = [float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_problems(a, first, count, K=None):
    """Trajectories [first, first+count) of the global C4 set (seed = 1 + global index)."""
    from fast_racing_b200 import synth
    K = a.samples if K is None else K
    return synth.config_problems("C4", count, seed0=1 + first, itg_spaces=K) if (a.pieces, a.faces) == (64, 12) else \
        synth.make_batch(1 + first, count, a.pieces, a.faces, K)


# ------------------------------------------------------------------------------------------------------------------
# CPU side (oracle port = the reference's algorithm restated, oracle/)
# ------------------------------------------------------------------------------------------------------------------
def cpu_baseline(problems_with_v, x_list, budget_s, cores, solve_tol=None, solve_count=0):
    """Oracle port on host cores: 1-core and all-core evals/sec on a bounded sample of the workload; optionally
    1-core and all-core solves/sec to the reference's stopping rule on the same trajectories."""
    from oracle import binding as ob
    nthr = cores["effective"]
    n_or = min(len(problems_with_v), max(nthr, 8))
    oracles = [ob.Oracle(p) for p in problems_with_v[:n_or]]
    t1 = oracles[0].time_evals(x_list[0], 3) / 3
    reps1 = max(3, int(0.25 * budget_s / max(t1, 1e-6)))
    one = oracles[0].time_evals(x_list[0], reps1)
    one_core = reps1 / one
    cal = ob.time_evals_mt(oracles, x_list[:n_or], 20, nthr) / 20.0  # calibrated under load
    reps = max(2, int(0.75 * budget_s / max(cal, 1e-6)))
    secs = ob.time_evals_mt(oracles, x_list[:n_or], reps, nthr)
    allc = n_or * reps / secs
    out = {"value": allc, "unit": "evals/s", "cores": nthr, "kind": "port",
           "sample": f"{n_or} trajectories of the same workload x {reps} evaluations each on {nthr} threads (one independent optimiser per thread); "
                     f"1 core: {one_core:.1f} evals/s over {reps1} evaluations",
           "one_core_value": one_core, "parallel_speedup": allc / one_core, "host": cores}
    if solve_tol is not None and solve_count > 0:
        for o in oracles:
            o.set_max_iterations(200000)  # safety net only (the reference has none): ~50x the iterations these solves need
        # full solves to the reference's stopping rule (lbfgs.hpp:1327-1345): one trajectory on one thread, then one per thread
        s1, r1 = ob.time_solves_mt(oracles[:1], solve_tol, 1)
        ns = min(n_or, max(nthr, 1), solve_count)
        sa, ra = ob.time_solves_mt(oracles[:ns], solve_tol, nthr)
        out["solves"] = {"one_core_value": 1.0 / s1, "value": ns / sa, "unit": "solves/s", "cores": nthr, "rel_tol": solve_tol,
                         "sample": f"1 solve on 1 thread ({s1:.2f} s, {int(r1['iterations'][0])} iterations, {int(r1['evaluations'][0])} evaluations); "
                                   f"{ns} solves on {nthr} threads ({sa:.2f} s)",
                         "iterations_mean": float(ra["iterations"].mean()), "evaluations_mean": float(ra["evaluations"].mean())}
    for o in oracles:
        o.close()
    return out


def run_reference(a):
    """--impl reference: the reference's CPU implementation of the path (oracle port, see DESIGN.md: the
    reference itself needs Eigen/ROS and cannot be built here) on all usable host cores, same config/metric.
    Nothing of the product is loaded here: the V-polytopes come from the oracle's own enumeration."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding as ob
    cores = effective_cores()
    nthr = cores["effective"]
    n_or = max(nthr, 8)
    probs = build_problems(a, 0, n_or)
    rng = np.random.default_rng(99)
    oracles, xs = [], []
    for p in probs:
        off, verts = ob.vpolys(p)
        o = ob.Oracle(p.with_vpolys(off, verts))
        x = o.initial_guess()
        xs.append(x + a.noise * rng.normal(size=x.size))
        oracles.append(o)
    # calibrate UNDER LOAD (all threads busy): per-evaluation time rises with SMT / memory contention
    cal = ob.time_evals_mt(oracles, xs, 20, nthr) / 20.0  # seconds for one evaluation of every trajectory on all threads
    # second stage: ~3 s sustained (short bursts run several times faster per evaluation than a saturated box)
    r2 = max(20, int(min(3.0, a.ref_seconds / 10.0) / max(cal, 1e-6)))
    cal = max(cal, ob.time_evals_mt(oracles, xs, r2, nthr) / r2)
    one = oracles[0].time_evals(xs[0], max(3, int(1.0 / max(cal, 1e-6) / 4)))
    one_core = max(3, int(1.0 / max(cal, 1e-6) / 4)) / one
    per_step_budget = a.ref_seconds / max(1, a.steps + a.warmup)
    reps = max(1, int(per_step_budget / max(cal, 1e-6)))
    for _ in range(a.warmup):
        ob.time_evals_mt(oracles, xs, reps, nthr)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        ob.time_evals_mt(oracles, xs, reps, nthr)
    dt = time.perf_counter() - t0
    evals = n_or * reps * a.steps
    v = evals / dt
    line = {
        "impl": "reference", "metric": "L-BFGS cost+grad evals/sec", "value": v, "unit": "evals/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": f"{n_or} trajectories x {reps} evaluations per step"},
        "cpu_baseline": {"value": v, "unit": "evals/s", "cores": nthr, "kind": "port",
                         "sample": f"{n_or} trajectories x {reps} evaluations per step on {nthr} threads",
                         "one_core_value": one_core, "parallel_speedup": v / one_core, "host": cores},
        "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "points_per_sec": v * a.pieces * (a.samples + 1),
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------------------------
class Ctx:
    """torch / distributed plumbing shared by the legs."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.dev = torch.device("cuda", self.local_rank)
        self.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=self.dev)  # 256 MiB > 126 MB L2
        self.stream = torch.cuda.current_stream()
        self.sp = self.stream.cuda_stream

    def sync_all(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def argmin_gather(self, cost, gidx):
        """Cross-rank arg-min of (cost, global index): one all_gather of 16 B per rank (NCCL), ties -> lowest index.
        Returns (cost, index, gather_ms); gather_ms is None on one rank."""
        torch, dist = self.torch, self.dist
        if self.world == 1:
            return cost, gidx, None
        mine = torch.tensor([cost, float(gidx)], dtype=torch.float64, device=self.dev)
        allv = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(allv, mine)  # warm-up: the first collective of a shape pays NCCL's lazy channel set-up
        self.sync_all()
        g0 = time.perf_counter()
        dist.all_gather(allv, mine)
        torch.cuda.synchronize()
        gather_ms = (time.perf_counter() - g0) * 1e3
        arr = torch.stack(allv).cpu().numpy()
        j = int(np.lexsort((arr[:, 1], arr[:, 0]))[0])
        return float(arr[j, 0]), int(arr[j, 1]), gather_ms


def timed_evals(cx, batch, x_host, steps, warmup, per_kernel=True):
    """Device-resident evaluation timing: `steps` evaluations, L2 flushed between them (outside the CUDA-event pairs).
    Returns (total_ms max over ranks, kernels_ms[3] of this rank)."""
    torch = cx.torch
    xd = torch.from_numpy(x_host).to(cx.dev)
    gd = torch.zeros_like(xd)
    cd = torch.zeros(batch.count, dtype=torch.float64, device=cx.dev)
    for _ in range(max(warmup, 3)):
        batch.eval_device(xd.data_ptr(), cd.data_ptr(), gd.data_ptr(), cx.sp)
    cx.sync_all()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        cx.flush.fill_(float(k))
        ev[k][0].record(cx.stream)
        batch.eval_device(xd.data_ptr(), cd.data_ptr(), gd.data_ptr(), cx.sp)
        ev[k][1].record(cx.stream)
    cx.sync_all()
    total_ms = float(sum(e0.elapsed_time(e1) for e0, e1 in ev))
    kms = np.zeros(3)
    if per_kernel:
        for k in range(steps):
            cx.flush.fill_(float(k))
            kms += batch.profile_eval(xd.data_ptr(), cd.data_ptr(), gd.data_ptr(), cx.sp, 1) / steps
        torch.cuda.synchronize()
    assert bool(torch.isfinite(cd).all()) and bool(torch.isfinite(gd).all())
    return cx.max_over_ranks([total_ms])[0], kms, (xd, gd, cd)


def penalty_rooflines(kms_pen_ms, B, M, S, F, V, n_mean, hbm_peak, fp64_peak, sass):
    """Both roofline fractions of the penalty kernel.  fp64: the algorithmic estimate and, when an ncu
    capture of this shape is committed under profiles/, the SASS-counted DFMA/DMUL/DADD thread instructions."""
    ab = algorithmic_bytes(M, F, V, n_mean)
    pen_s = kms_pen_ms * 1e-3
    hbm = ab["penalty_kernel"] * B / pen_s / 1e9
    fl = algorithmic_flops(M, S, F) * B
    out = {"hbm_GBps": hbm, "hbm_frac": hbm / hbm_peak,
           "fp64_algorithmic_TFLOPs": fl / pen_s / 1e12, "fp64_algorithmic_frac": (fl / pen_s / 1e12 / fp64_peak) if fp64_peak > 0 else None}
    if sass is not None:
        out["fp64_sass_TFLOPs"] = sass["flops"] / pen_s / 1e12
        out["fp64_sass_frac"] = (sass["flops"] / pen_s / 1e12 / fp64_peak) if fp64_peak > 0 else None
    return out, ab, fl


def load_profile_json(name):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name)))
    except Exception:
        return None


def solve_leg(cx, api, probs, tol, first_global, label, slots=0):
    """Batched L-BFGS solves of `probs` on this rank + the cross-rank arg-min.  Returns the record (rank-agnostic values).
    slots > 0: queue-fed solver slots (frx_batch_set_solver_slots) — `slots` trajectories in flight, the rest wait their turn."""
    t_c = time.perf_counter()
    sbatch = api.Batch(probs, device=cx.local_rank)
    create_ms = (time.perf_counter() - t_c) * 1e3  # V-polytope extraction (device), tables, packing, uploads
    sb = sbatch.count
    sbatch.set_solver_slots(slots)
    cx.sync_all()
    t0 = time.perf_counter()
    r = sbatch.optimize(tol)
    wall = time.perf_counter() - t0
    bi, bc = sbatch.argmin()
    best_c, best_i, gather_ms = cx.argmin_gather(bc, first_global + bi)
    wall_max = cx.max_over_ranks([wall])[0]
    rec = {"workload": label, "value": sb * cx.world / wall_max, "unit": "solves/s", "batch_per_gpu": sb,
           "solver_slots_per_gpu": slots if 0 < slots < sb else sb, "rel_tol": tol,
           "batch_create_ms": create_ms,
           "wall_s_max_over_ranks": wall_max, "device_ms": r["device_ms"], "ticks": int(r["ticks"]),
           "iterations_mean": float(r["iterations"].mean()), "evaluations_mean": float(r["evaluations"].mean()),
           "evaluations_max": int(r["evaluations"].max()),
           # line-search error exits (the reference swallows them too, CPU.hpp:1249): the iterate returned is the last accepted one
           "status_ok_frac": float((r["status"] >= 0).mean()),
           "argmin": {"cost": best_c, "global_index": best_i}, "nccl_gather_ms": gather_ms,
           # lockstep occupancy: share of (trajectory, tick) slots that still evaluated
           "active_slot_frac": float(r["evaluations"].mean()) / max(int(r["ticks"]), 1)}
    uk = r["update_kernel"]
    if uk["sampled_launches"] > 0 and uk["device_ms"] > 0:
        # the HBM-streaming kernel of the path: history bytes of the sampled launches / their device time
        rec["update_kernel"] = {"sampled_launches": uk["sampled_launches"], "mean_us": 1e3 * uk["device_ms"] / uk["sampled_launches"],
                                "history_GB": uk["history_bytes"] / 1e9,
                                "achieved_GBps": uk["history_bytes"] / (uk["device_ms"] * 1e-3) / 1e9}
    sbatch.close()
    return rec


def corridor_leg(cx, api, synth, n_points=200000, n_seg=64):
    """Row f-4: the segment step of the corridor generation, batched on the device, beside the host path (one thread)."""
    rng = np.random.default_rng(3)
    centres = synth.corridor_centres(rng, n_seg)
    obs = synth.obstacle_cloud(3, centres, n_points, margin=8.0)
    segs = np.array([np.r_[a, b] for a, b in zip(centres[:-1], centres[1:])])
    bbox = np.array([4.0, 4.0, 2.5])
    cloud = api.Cloud(obs, device=cx.local_rank)
    cloud.dilate_batch(segs[:2], bbox)
    t0 = time.perf_counter()
    _, cnt, _ = cloud.dilate_batch(segs, bbox)
    dev_ms = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    for sg in segs[:8]:
        api.corridor_dilate(obs, sg[:3], sg[3:], bbox)
    host_ms = (time.perf_counter() - t0) * 1e3 / 8
    cloud.close()
    return {"workload": f"{len(segs)} path segments against an obstacle cloud of {len(obs)} points (LineSegment3D::dilate each)",
            "device_batch_ms": dev_ms, "device_us_per_segment": dev_ms * 1e3 / len(segs), "host_ms_per_segment_1_thread": host_ms,
            "faces_per_segment_mean": float(cnt.mean()), "note": "device time includes the H2D of the segments and the D2H of the faces"}


def single_solve_latency(cx, api, name, tol):
    """C1 / C2: ONE trajectory, the reference's own call (MinCoPlan_CPU.cpp:114-124): GPU solve latency beside the CPU port's."""
    from fast_racing_b200 import synth
    from oracle import binding as ob
    p = synth.config_problems(name, 1)[0]
    b = api.Batch([p], device=cx.local_rank)
    b.optimize(tol)  # warm-up (solver allocation, kernel configuration)
    t0 = time.perf_counter()
    r = b.optimize(tol)
    gpu_s = time.perf_counter() - t0
    off, verts = b.vpolys(0)
    o = ob.Oracle(p.with_vpolys(off, verts))
    o.set_max_iterations(200000)  # safety net only
    t0 = time.perf_counter()
    ro = o.optimize(tol)
    cpu_s = time.perf_counter() - t0
    d = b.dims(0)
    rec = {"pieces": d["pieces"], "lbfgs_dim": d["n"], "rel_tol": tol, "gpu_solve_ms": gpu_s * 1e3, "gpu_device_ms": r["device_ms"],
           "gpu_evaluations": int(r["evaluations"][0]), "gpu_final_cost": float(r["final_cost"][0]), "gpu_status": int(r["status"][0]),
           "cpu_solve_ms": cpu_s * 1e3, "cpu_evaluations": int(ro["evaluations"]), "cpu_final_cost": float(ro["final_cost"]),
           "cpu_kind": "port, 1 thread", "gpu_us_per_evaluation": gpu_s * 1e6 / max(int(r["evaluations"][0]), 1),
           "cpu_us_per_evaluation": cpu_s * 1e6 / max(int(ro["evaluations"]), 1)}
    o.close()
    b.close()
    return rec


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
        return
    cx = Ctx()
    torch = cx.torch
    from fast_racing_b200 import api, synth
    rank, world = cx.rank, cx.world
    # one rank per GPU: this rank's threads and host allocations go to the GPU's own NUMA node (frx.h); the CPU-baseline leg
    # gets the process's ORIGINAL affinity mask back (it must see every core the box gives this job)
    host_cores = effective_cores()
    orig_affinity = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    numa_node = api.load_library().frx_bind_thread_to_device(cx.local_rank)

    probs = build_problems(a, rank * a.batch, a.batch)  # rank r owns trajectories [r*B, (r+1)*B) of the global batch
    batch = api.Batch(probs, device=cx.local_rank)
    B = batch.count
    d0 = batch.dims(0)
    x0 = batch.initial_guess()
    rng = np.random.default_rng(1234 + rank)
    x_host = x0 + a.noise * rng.normal(size=x0.size)

    # clocks are sampled (nvidia-smi, ~10 Hz) from here to the end of the end-to-end leg: warm-up, the timed region, the
    # per-kernel pass and the host-buffer pass all run the same three kernels back to back
    sampler = ClockSampler(cx.local_rank)
    sampler.start()
    total_ms, kms, (xd, gd, cd) = timed_evals(cx, batch, x_host, a.steps, a.warmup)

    # ---- end-to-end through the public host-buffer API (H2D of x, D2H of cost + grad every step) ----
    # The caller keeps x / reads (cost, grad) in the batch's pinned host buffers (frx_batch_host_buffers), so each step is
    # exactly: H2D copy of x from pinned memory, three kernels, D2H of cost + grad into pinned memory, synchronise.
    x_pin, cost_h, grad_h = batch.host_buffers()
    for _ in range(2):
        x_pin[:] = x_host
        batch.eval_pinned()
    cx.sync_all()
    t0 = time.perf_counter()
    for k in range(a.steps):
        x_pin[0] = x_host[0] + 1e-12 * k  # the step's input is (re)written on the host every step
        batch.eval_pinned()
        sink = float(cost_h[0]) + float(grad_h[-1])  # the step's result is read on the host
    e2e_s = time.perf_counter() - t0
    assert np.isfinite(cost_h).all() and np.isfinite(grad_h).all()
    # the timed regions are only milliseconds long: keep the same kernels running until the sampler has seen the GPU
    # under this load for >= 1.5 s (not part of any reported time)
    t_busy = time.perf_counter()
    while time.perf_counter() - t_busy < 1.5 or len(sampler.rows) < 8:
        for _ in range(50):
            batch.eval_device(xd.data_ptr(), cd.data_ptr(), gd.data_ptr(), cx.sp)
        torch.cuda.synchronize()
        if time.perf_counter() - t_busy > 6.0:
            break
    clocks = sampler.summary()
    e2e_ms = cx.max_over_ranks([e2e_s * 1e3])[0]
    evals = B * world * a.steps

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    fp64_peak = batch.lib.frx_measure_fp64_tflops()
    M, S, F = d0["pieces"], a.samples + 1, a.faces
    off0, _ = batch.vpolys(0)
    V = float(np.diff(off0).mean())
    n_mean = batch.total_x / B
    prof = load_profile_json("r02_ncu_eval.json")  # ncu --set full of this shape: DRAM traffic + SASS fp64 instruction counts
    same_shape = prof is not None and prof.get("shape") == [B, M, S, F]
    pen_prof = prof["kernels"].get("penalty_kernel") if same_shape else None

    # ---- C4 solves/sec leg: full batched L-BFGS solves of a sub-batch + the NCCL arg-min gather ----
    solves = None
    if a.solve_batch > 0:
        sb = min(a.solve_batch, B)
        solves = solve_leg(cx, api, probs[:sb], 1e-5, rank * a.batch, f"C4 sub-batch: {sb} trajectories/GPU x {M} pieces x K={a.samples}")

    extras = {}
    if not a.no_extras:
        if a.solve_batch > 0 and B > a.solve_batch:
            # the whole C4 batch solved at once (one slot per trajectory), and the same batch fed through half as many slots:
            # per-trajectory results are identical (tests/test_parity_gpu.py::test_solver_slots_do_not_change_results)
            extras["solves_full_batch"] = solve_leg(cx, api, probs, 1e-5, rank * a.batch,
                                                    f"C4: {B} trajectories/GPU x {M} pieces x K={a.samples}, one solver slot each")
            extras["solves_queue"] = solve_leg(cx, api, probs, 1e-5, rank * a.batch,
                                               f"C4: {B} trajectories/GPU x {M} pieces x K={a.samples} through {B // 2} solver slots",
                                               slots=B // 2)
        # ---- strong scaling: north_star's "1024-trajectory x 64-piece x 32-sample batch" in TOTAL over the N ranks ----
        total = a.batch
        per = total // world
        if world == 1:
            extras["strong"] = {"total_trajectories": total, "per_gpu": per, "value": evals / (total_ms * 1e-3), "unit": "evals/s",
                                "ms_per_step": total_ms / a.steps, "note": "N=1: identical to the headline measurement"}
        else:
            sprobs = build_problems(a, rank * per, per)
            sbt = api.Batch(sprobs, device=cx.local_rank)
            xs0 = sbt.initial_guess()
            xs = xs0 + a.noise * np.random.default_rng(4321 + rank).normal(size=xs0.size)
            s_ms, s_k, _ = timed_evals(cx, sbt, xs, a.steps, a.warmup)
            extras["strong"] = {"total_trajectories": per * world, "per_gpu": per, "value": per * world * a.steps / (s_ms * 1e-3),
                                "unit": "evals/s", "ms_per_step": s_ms / a.steps,
                                "kernels_ms": {"minco_forward": s_k[0], "penalty": s_k[1], "minco_backward": s_k[2]}}
            sbt.close()
        # ---- K sweep (config 4): same corridors and V-polytopes, K in {8..128}; weak scaling like the headline ----
        sweep = []
        vp = [batch.vpolys(i) for i in range(B)]
        for K in SWEEP_K:
            if K == a.samples:
                k_ms, k_k, Sk = total_ms, kms, S
            else:
                pk = [api.Problem(p.ini, p.fin, p.face_off, p.faces, dict(p.params, itg_spaces=K), vp[i][0], vp[i][1]) for i, p in enumerate(probs)]
                bk = api.Batch(pk, device=cx.local_rank)
                k_ms, k_k, _ = timed_evals(cx, bk, x_host, max(5, a.steps // 2), 3)
                k_ms *= a.steps / max(5, a.steps // 2)
                bk.close()
                Sk = K + 1
            rf, _, _ = penalty_rooflines(k_k[1], B, M, Sk, F, V, n_mean, hbm_peak, fp64_peak, None)
            v = evals / (k_ms * 1e-3)
            sweep.append({"K": K, "value": v, "unit": "evals/s", "points_per_sec": v * M * Sk, "ms_per_step": k_ms / a.steps,
                          "kernels_ms": {"minco_forward": k_k[0], "penalty": k_k[1], "minco_backward": k_k[2]},
                          "penalty_hbm_frac": rf["hbm_frac"], "penalty_fp64_algorithmic_frac": rf["fp64_algorithmic_frac"]})
        extras["sweep"] = sweep
        del vp
        # ---- the other BASELINE configurations ----
        cfgs = {}
        if a.c5_batch > 0:  # C5: Monte-Carlo gate perturbations of one moderncity-like track, sharded, arg-min over all ranks
            c5 = synth.config_problems("C5", a.c5_batch, seed0=1 + rank * a.c5_batch)
            cfgs["C5"] = solve_leg(cx, api, c5, 1e-5, rank * a.c5_batch,
                                   f"C5: {a.c5_batch} perturbed tracks/GPU x 24 pieces x K=48, moderncity parameters, OptRelTol 1e-5")
            del c5
        if world == 1:  # single-GPU configurations
            c3 = synth.config_problems("C3", 64)
            b3 = api.Batch(c3, device=cx.local_rank)
            x3 = b3.initial_guess()
            x3 = x3 + a.noise * np.random.default_rng(77).normal(size=x3.size)
            m3, k3, _ = timed_evals(cx, b3, x3, a.steps, a.warmup)
            cfgs["C3"] = {"workload": "C3: 64 trajectories x 32 pieces x K=16", "value": 64 * a.steps / (m3 * 1e-3), "unit": "evals/s",
                          "ms_per_step": m3 / a.steps, "kernels_ms": {"minco_forward": k3[0], "penalty": k3[1], "minco_backward": k3[2]}}
            b3.close()
            if not a.no_cpu_baseline:
                cfgs["C1"] = single_solve_latency(cx, api, "C1", 1e-6)
                cfgs["C2"] = single_solve_latency(cx, api, "C2", 1e-6)
            cfgs["corridor"] = corridor_leg(cx, api, synth)
        extras["configs"] = cfgs

    if rank == 0:
        rf, ab, fl = penalty_rooflines(kms[1], B, M, S, F, V, n_mean, hbm_peak, fp64_peak, pen_prof)
        value = evals / (total_ms * 1e-3)
        line = {
            "metric": "L-BFGS cost+grad evals/sec", "value": value, "unit": "evals/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "batch_per_gpu": B, "pieces": M, "samples_per_piece": S, "faces": F,
                       "lbfgs_dim_mean": n_mean, "eval_point": f"setInitial+backward start + N(0,{a.noise}^2)",
                       "l2": "flushed between timed steps (256 MiB write outside the event pairs)",
                       "sharding": f"batch split over {world} rank(s), no data-path collective"},
            "points_per_sec": value * M * S,
            "e2e": {"value": evals / (e2e_ms * 1e-3), "unit": "evals/s", "h2d_bytes_per_step": int(batch.total_x * 8),
                    "d2h_bytes_per_step": int((batch.total_x + B) * 8), "ms_per_step": e2e_ms / a.steps,
                    "api": "frx_batch_eval on the batch's pinned host buffers (host x -> host cost, grad; H2D + D2H inside the timed region)",
                    "numa_node_rank0": numa_node},
            "gpu_launches": int(batch.eval_launches() * a.steps),
            "kernels_ms": {"minco_forward": kms[0], "penalty": kms[1], "minco_backward": kms[2]},
            "roofline": {"kernel": "penalty_kernel", "bound": "hbm", "achieved": rf["hbm_GBps"], "peak": hbm_peak, "unit": "GB/s",
                         "frac": rf["hbm_frac"], "traffic": pen_prof["traffic_bytes"] if pen_prof else None,
                         "traffic_source": "profiles/r02_ncu_eval.json (ncu --set full, same workload)" if pen_prof else None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": ab["penalty_kernel"] * B,
                         "note": "fp64-pipe-bound kernel: see roofline_fp64 (SURVEY.md §8d caveat)"},
            "roofline_fp64": {"kernel": "penalty_kernel", "bound": "fp64", "achieved": rf["fp64_algorithmic_TFLOPs"], "peak": fp64_peak,
                              "unit": "TFLOP/s", "frac": rf["fp64_algorithmic_frac"],
                              "peak_source": "DFMA micro-benchmark in libfrx.so, measured in this run",
                              "algorithmic_flops_per_launch": fl,
                              "sass": ({"flops_per_launch": pen_prof["flops"], "dfma": pen_prof["dfma"], "dmul": pen_prof["dmul"],
                                        "dadd": pen_prof["dadd"], "achieved": rf["fp64_sass_TFLOPs"], "frac": rf["fp64_sass_frac"],
                                        "source": "profiles/r02_ncu_eval.json: smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on.sum "
                                                  "of one launch of this shape; flops = 2 dfma + dmul + dadd"} if pen_prof else None)},
            "pipeline_hbm": {"achieved": ab["pipeline"] * B / (total_ms / a.steps * 1e-3) / 1e9, "unit": "GB/s",
                             "algorithmic_bytes_per_step": ab["pipeline"] * B,
                             "dram_traffic_per_step": (sum(k["traffic_bytes"] for k in prof["kernels"].values()) if same_shape else None)},
            "clocks": clocks,
            "penalty_launch": batch.penalty_launch(),
        }
        if solves is not None:
            line["solves"] = solves
        line.update(extras)
        if not a.no_cpu_baseline:
            if orig_affinity is not None:
                os.sched_setaffinity(0, orig_affinity)
            cores = host_cores
            pv, xs = [], []
            for i in range(min(B, max(cores["effective"], 8))):
                off, verts = batch.vpolys(i)
                pv.append(probs[i].with_vpolys(off, verts))
                xs.append(x_host[batch.x_off[i]:batch.x_off[i + 1]].copy())
            line["cpu_baseline"] = cpu_baseline(pv, xs, a.cpu_seconds, cores, solve_tol=1e-5 if a.solve_batch > 0 else None,
                                                solve_count=a.solve_batch)
        print(json.dumps(line), flush=True)
    batch.close()
    if world > 1:
        cx.dist.barrier()
        cx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
