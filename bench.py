#!/usr/bin/env python
"""Benchmark of the hot path: batched L-BFGS cost+gradient evaluations (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (libfrx.so)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

A "step" = one batched cost+grad evaluation of every trajectory of the workload (one call of the
lbfgs_evaluate_t boundary per trajectory).  Workload = BASELINE.json configs[3] ("C4"): per GPU 1024
candidate trajectories x 64 pieces x 12-face polytopes x 32 quadrature intervals (33 samples) per piece,
evaluated at the optimiser's starting point plus N(0, 0.5^2) noise so that every penalty family (P,V,T,B) is active.
Multi-GPU: the batch shards across ranks (weak scaling: 1024 trajectories per GPU), no data-path
collective; one NCCL all-gather of (cost, index) per solve for the arg-min (timed separately, reported).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


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
    ap.add_argument("--solve-batch", type=int, default=256, help="trajectories of the solves/sec leg (0 = skip)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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


def build_problems(a, rank):
    from fast_racing_b200 import synth
    # distinct seeds per rank: rank r owns trajectories [r*B, (r+1)*B) of the global batch
    return synth.config_problems("C4", a.batch, seed0=1 + rank * a.batch, itg_spaces=a.samples) if (a.pieces, a.faces) == (64, 12) else \
        synth.make_batch(1 + rank * a.batch, a.batch, a.pieces, a.faces, a.samples)


def cpu_baseline(problems_with_v, x_list, a, budget_s):
    """Oracle port on host cores: 1-core and all-core evals/sec on a bounded sample of the workload."""
    from oracle import binding as ob
    ncpu = os.cpu_count() or 1
    n_or = min(len(problems_with_v), max(ncpu, 8))
    oracles = [ob.Oracle(p) for p in problems_with_v[:n_or]]
    t1 = oracles[0].time_evals(x_list[0], 3) / 3
    reps1 = max(3, int(0.25 * budget_s / max(t1, 1e-6)))
    one = oracles[0].time_evals(x_list[0], reps1)
    one_core = reps1 / one
    cal = ob.time_evals_mt(oracles, x_list[:n_or], 20, ncpu) / 20.0  # calibrated under load
    reps = max(2, int(0.75 * budget_s / max(cal, 1e-6)))
    secs = ob.time_evals_mt(oracles, x_list[:n_or], reps, ncpu)
    allc = n_or * reps / secs
    for o in oracles:
        o.close()
    return {"value": allc, "unit": "evals/s", "cores": ncpu, "kind": "port",
            "sample": f"{n_or} trajectories of the same workload x {reps} evaluations each on {ncpu} threads (one independent optimiser per thread); "
                      f"1 core: {one_core:.1f} evals/s over {reps1} evaluations",
            "one_core_value": one_core}


def run_reference(a):
    """--impl reference: the reference's CPU implementation of the path (oracle port, see DESIGN.md: the
    reference itself needs Eigen/ROS and cannot be built here) on all host cores, same config/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from fast_racing_b200 import api
    from oracle import binding as ob
    ncpu = os.cpu_count() or 1
    n_or = max(ncpu, 8)
    sub = argparse.Namespace(**vars(a))
    sub.batch = n_or
    probs = build_problems(sub, 0)
    rng = np.random.default_rng(99)
    oracles, xs = [], []
    for p in probs:
        off, verts = api.host_vpolys(p)  # host-only geometry (no GPU), shared setup code
        o = ob.Oracle(p.with_vpolys(off, verts))
        x = o.initial_guess()
        xs.append(x + a.noise * rng.normal(size=x.size))
        oracles.append(o)
    # calibrate UNDER LOAD (all threads busy): per-evaluation time rises with SMT / memory contention
    cal = ob.time_evals_mt(oracles, xs, 20, ncpu) / 20.0  # seconds for one evaluation of every trajectory on all threads
    # second stage: ~3 s sustained (short bursts run several times faster per evaluation than a saturated box)
    r2 = max(20, int(min(3.0, a.ref_seconds / 10.0) / max(cal, 1e-6)))
    cal = max(cal, ob.time_evals_mt(oracles, xs, r2, ncpu) / r2)
    per_step_budget = a.ref_seconds / max(1, a.steps + a.warmup)
    reps = max(1, int(per_step_budget / max(cal, 1e-6)))
    for _ in range(a.warmup):
        ob.time_evals_mt(oracles, xs, reps, ncpu)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        ob.time_evals_mt(oracles, xs, reps, ncpu)
    dt = time.perf_counter() - t0
    evals = n_or * reps * a.steps
    v = evals / dt
    line = {
        "impl": "reference", "metric": "L-BFGS cost+grad evals/sec", "value": v, "unit": "evals/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": f"{n_or} trajectories x {reps} evaluations per step"},
        "cpu_baseline": {"value": v, "unit": "evals/s", "cores": ncpu, "kind": "port",
                         "sample": f"{n_or} trajectories x {reps} evaluations per step on {ncpu} threads"},
        "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "points_per_sec": v * a.pieces * (a.samples + 1),
    }
    print(json.dumps(line), flush=True)


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
        return
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from fast_racing_b200 import api

    probs = build_problems(a, rank)
    batch = api.Batch(probs, device=local_rank)
    B = batch.count
    d0 = batch.dims(0)
    x0 = batch.initial_guess()
    rng = np.random.default_rng(1234 + rank)
    x_host = x0 + a.noise * rng.normal(size=x0.size)

    dev = torch.device("cuda", local_rank)
    xd = torch.from_numpy(x_host).to(dev)
    gd = torch.zeros_like(xd)
    cd = torch.zeros(B, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MiB > 126 MB L2
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # clocks are sampled (nvidia-smi, ~10 Hz) from here to the end of the end-to-end leg: warm-up, the timed region, the
    # per-kernel pass and the host-buffer pass all run the same three kernels back to back
    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- warm-up ----
    for _ in range(max(a.warmup, 3)):
        batch.eval_device(xd.data_ptr(), cd.data_ptr(), gd.data_ptr(), sp)
    sync_all()

    # ---- timed region: K steps, L2 flushed between steps (outside the event pairs) ----
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    sync_all()
    for k in range(a.steps):
        flush.fill_(float(k))
        ev[k][0].record(stream)
        batch.eval_device(xd.data_ptr(), cd.data_ptr(), gd.data_ptr(), sp)
        ev[k][1].record(stream)
    sync_all()
    step_ms = np.array([e0.elapsed_time(e1) for e0, e1 in ev])
    total_ms = float(step_ms.sum())
    # per-kernel times (same flush discipline), for the roofline
    kms = np.zeros(3)
    for k in range(a.steps):
        flush.fill_(float(k))
        kms += batch.profile_eval(xd.data_ptr(), cd.data_ptr(), gd.data_ptr(), sp, 1) / a.steps
    torch.cuda.synchronize()

    # ---- end-to-end through the public host-buffer API (H2D of x, D2H of cost + grad every step) ----
    # The caller keeps x / reads (cost, grad) in the batch's pinned host buffers (frx_batch_host_buffers), so each step is
    # exactly: H2D copy of x from pinned memory, three kernels, D2H of cost + grad into pinned memory, synchronise.
    x_pin, cost_h, grad_h = batch.host_buffers()
    for _ in range(2):
        x_pin[:] = x_host
        batch.eval_pinned()
    sync_all()
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
            batch.eval_device(xd.data_ptr(), cd.data_ptr(), gd.data_ptr(), sp)
        torch.cuda.synchronize()
        if time.perf_counter() - t_busy > 6.0:
            break
    clocks = sampler.summary()

    # max over ranks
    t = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])
    evals = B * world * a.steps

    # ---- solves/sec leg: full batched L-BFGS solves of a sub-batch + the NCCL arg-min gather ----
    solves = None
    if a.solve_batch > 0:
        sb = min(a.solve_batch, B)
        sbatch = api.Batch(probs[:sb], device=local_rank)
        tol = 1e-5
        t0 = time.perf_counter()
        r = sbatch.optimize(tol)
        wall = time.perf_counter() - t0
        bi, bc = sbatch.argmin()
        gather_ms = None
        best = (bc, rank * a.batch + bi)
        if world > 1:
            mine = torch.tensor([bc, float(rank * a.batch + bi)], dtype=torch.float64, device=dev)
            allv = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allv, mine)  # warm-up: the first collective of a shape pays NCCL's lazy channel set-up
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            g0 = time.perf_counter()
            dist.all_gather(allv, mine)
            torch.cuda.synchronize()
            gather_ms = (time.perf_counter() - g0) * 1e3
            arr = torch.stack(allv).cpu().numpy()
            j = int(np.lexsort((arr[:, 1], arr[:, 0]))[0])
            best = (float(arr[j, 0]), int(arr[j, 1]))
        tt = torch.tensor([wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        solves = {"value": sb * world / float(tt[0]), "unit": "solves/s", "batch_per_gpu": sb, "rel_tol": tol,
                  "device_ms": r["device_ms"], "ticks": int(r["ticks"]),
                  "iterations_mean": float(r["iterations"].mean()), "evaluations_mean": float(r["evaluations"].mean()),
                  "status_ok_frac": float((r["status"] >= 0).mean()), "argmin": {"cost": best[0], "global_index": best[1]},
                  "nccl_gather_ms": gather_ms,
                  # lockstep occupancy: share of (trajectory, tick) slots that still evaluated
                  "active_slot_frac": float(r["evaluations"].mean()) / max(int(r["ticks"]), 1)}
        uk = r["update_kernel"]
        if uk["sampled_launches"] > 0 and uk["device_ms"] > 0:
            # the HBM-streaming kernel of the path: history bytes of the sampled launches / their device time
            solves["update_kernel"] = {"sampled_launches": uk["sampled_launches"], "mean_us": 1e3 * uk["device_ms"] / uk["sampled_launches"],
                                       "history_GB": uk["history_bytes"] / 1e9,
                                       "achieved_GBps": uk["history_bytes"] / (uk["device_ms"] * 1e-3) / 1e9}
        sbatch.close()

    if rank == 0:
        M, S, F = d0["pieces"], a.samples + 1, a.faces
        off, _ = batch.vpolys(0)
        V = float(np.diff(off).mean())
        n_mean = batch.total_x / B
        ab = algorithmic_bytes(M, F, V, n_mean)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        traffic, traffic_src = None, None
        try:  # dram__bytes_read+write of the penalty kernel per launch, from the committed `ncu --set full` capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01d_dram_traffic.json")))
            if (B, M, S, F) == (1024, 64, 33, 12):
                traffic = tj["kernels"]["penalty_kernel"]["traffic_bytes"]
                traffic_src = "profiles/r01d_dram_traffic.json (ncu --set full, same workload)"
        except Exception:
            pass
        pen_s = kms[1] * 1e-3
        achieved = ab["penalty_kernel"] * B / pen_s / 1e9
        fp64_peak = batch.lib.frx_measure_fp64_tflops()
        fl = algorithmic_flops(M, S, F) * B
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
                    "api": "frx_batch_eval on the batch's pinned host buffers (host x -> host cost, grad; H2D + D2H inside the timed region)"},
            "gpu_launches": int(batch.eval_launches() * a.steps),
            "kernels_ms": {"minco_forward": kms[0], "penalty": kms[1], "minco_backward": kms[2]},
            "roofline": {"kernel": "penalty_kernel", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": ab["penalty_kernel"] * B,
                         "note": "fp64-pipe-bound kernel: see roofline_fp64 (SURVEY.md §8d caveat)"},
            "roofline_fp64": {"kernel": "penalty_kernel", "bound": "fp64", "achieved": fl / pen_s / 1e12, "peak": fp64_peak,
                              "unit": "TFLOP/s", "frac": fl / pen_s / 1e12 / fp64_peak if fp64_peak > 0 else None,
                              "peak_source": "DFMA micro-benchmark in libfrx.so, measured in this run",
                              "algorithmic_flops_per_launch": fl},
            "pipeline_hbm": {"achieved": ab["pipeline"] * B / (total_ms / a.steps * 1e-3) / 1e9, "unit": "GB/s",
                             "algorithmic_bytes_per_step": ab["pipeline"] * B},
            "clocks": clocks,
            "penalty_launch": batch.penalty_launch(),
        }
        if solves is not None:
            line["solves"] = solves
        if not a.no_cpu_baseline:
            from oracle import binding as ob  # checker / baseline only
            pv, xs = [], []
            for i in range(min(B, max(os.cpu_count() or 1, 8))):
                off, verts = batch.vpolys(i)
                pv.append(probs[i].with_vpolys(off, verts))
                xs.append(x_host[batch.x_off[i]:batch.x_off[i + 1]].copy())
            line["cpu_baseline"] = cpu_baseline(pv, xs, a, a.cpu_seconds)
        print(json.dumps(line), flush=True)
    batch.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
