"""Turns an ncu report (--set full) into the tracked summaries under profiles/:
     <tag>_ncu_full_eval.csv   one column per evaluation kernel, the metrics the round notes quote
     <tag>_dram_traffic.json   dram bytes read + written per launch
     <tag>_ncu_eval.json       the same plus the SASS-counted fp64 thread instructions (dfma / dmul / dadd, flops = 2 dfma + dmul +
                               dadd) and the workload shape [B, M, S, F] (-> roofline.traffic and roofline_fp64.sass in bench.py)
   usage: python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r02 [B M S F] [free text about the capture]"""
import csv, io, json, subprocess, sys

rep, tag = sys.argv[1], sys.argv[2]
shape = [int(v) for v in sys.argv[3:7]] if len(sys.argv) >= 7 and all(v.isdigit() for v in sys.argv[3:7]) else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
WANT = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
ki = hdr.index("Kernel Name")
short = lambda n: n.split("(")[0].split("::")[-1]
seen, cols = set(), []
for r in data:  # first launch of each kernel
    if short(r[ki]) not in seen:
        seen.add(short(r[ki])); cols.append(r)
with open(tag + "_ncu_full_eval.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [short(r[ki]) for r in cols])
    for m in WANT:
        if m in hdr:
            i = hdr.index(m)
            w.writerow([m, units[i]] + [r[i] for r in cols])
mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
tm = {"ns": 1e-3, "us": 1.0, "ms": 1e3}
out = {"source": "ncu --set full --clock-control none, " + " ".join(sys.argv[3:]) + ", one launch each", "kernels": {}}
for r in cols:
    rd, wr = float(r[ir]) * mult[units[ir]], float(r[iw]) * mult[units[iw]]
    out["kernels"][short(r[ki])] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "duration_us_under_ncu": float(r[it]) * tm[units[it]],
                                    "traffic_bytes": rd + wr}
json.dump(out, open(tag + "_dram_traffic.json", "w"), indent=1)
ops = {}
for op in ("dfma", "dmul", "dadd"):
    name = f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed"
    ops[op] = hdr.index(name) if name in hdr else None
icyc = hdr.index("smsp__cycles_elapsed.avg") if "smsp__cycles_elapsed.avg" in hdr else None
if icyc is not None and all(v is not None for v in ops.values()):
    for r in cols:  # thread instructions of the launch = (sum over SMSPs per elapsed cycle) x elapsed cycles
        k = out["kernels"][short(r[ki])]
        cyc = float(r[icyc].replace(",", ""))
        for op, i in ops.items():
            k[op] = float(r[i].replace(",", "")) * cyc
        k["flops"] = 2.0 * k["dfma"] + k["dmul"] + k["dadd"]
        k["cycles_elapsed"] = cyc
    out["shape"] = shape
    json.dump(out, open(tag + "_ncu_eval.json", "w"), indent=1)
