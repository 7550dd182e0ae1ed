#!/usr/bin/env python
"""Turn the scratch output of tools/gpu_round.sh (gpurun_out/<tag>/) into the committed evidence under profiles/:
     profiles/<tag>_launches.md     per-kernel share of the ncu launch list (cold-cache, serialised: compare shares)
     profiles/<tag>_ncu_full.md     selected metrics of each `ncu --set full` capture (+ top stall reasons per source line)
     profiles/<tag>_bench_*.json    the bench lines of the same call
     profiles/traffic.json          DRAM bytes per point per solver (read by bench.py for roofline.traffic)
   python tools/summarise_ncu.py <tag>"""
import collections
import csv
import glob
import io
import json
import os
import re
import shutil
import subprocess
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sass_operands  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
src = os.path.join(ROOT, "gpurun_out", tag)
dst = os.path.join(ROOT, "profiles")
os.makedirs(dst, exist_ok=True)

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
] + ["smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % k for k in (
    "wait", "long_scoreboard", "short_scoreboard", "math_pipe_throttle", "barrier", "not_selected", "branch_resolving",
    "lg_throttle", "mio_throttle", "dispatch_stall", "no_instruction", "drain")]


def ncu_csv(rep, page):
    """Page of a capture: the CSV exported on the GPU box (<name>.<page>.csv) or, if the report came back, ncu -i."""
    pre = rep[:-len(".ncu-rep")] + "." + page + ".csv"
    if os.path.isfile(pre):
        out = open(pre).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    lines = [ln for ln in out.splitlines() if ln.startswith('"')]
    return list(csv.reader(io.StringIO("\n".join(lines))))


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


# ---- launch list ------------------------------------------------------------------------------------------------
ll = os.path.join(src, "launches.csv")
if os.path.isfile(ll):
    rows = [r for r in csv.reader(open(ll)) if len(r) > 10]
    h = rows[0]
    ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
    agg = defaultdict(list)
    for r in rows[1:]:
        v = num(r[vi])
        if v is not None:
            agg[(r[ki].split("(")[0], r[gi])].append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(dst, tag + "_launches.md"), "w") as f:
        f.write("# ncu launch list, `python bench.py --steps 2 --warmup 1 --no-cpu-baseline` (%s)\n\n" % tag)
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` -- per-launch times are cold-cache and "
                "serialised; compare SHARES with bench.py's per_solver.share_of_step, not absolutes.  Grids of 39063 / 9766 "
                "blocks are the 10 M-point device-resident steps; the 8192 / 2048-block grids are the 2 Mi-point chunks of "
                "the host-mode (e2e) pipeline.\n\n")
        f.write("| kernel | grid | launches | mean us | share of all kernel time |\n|---|---|---|---|---|\n")
        for (k, g), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("| `%s` | %s | %d | %.1f | %.3f |\n" % (k.replace("void ", ""), g, len(v), sum(v) / len(v) / 1e3, sum(v) / tot))
        # share within the device-resident steps only: bench.py runs (warmup + steps) = 3 device-resident steps first
        # (launches per step taken from the bench line of the same call), the host-mode (e2e) chunks follow
        per_step_launches = 8
        bj = os.path.join(src, "bench_10M.json")
        if os.path.isfile(bj):
            try:
                bd = json.load(open(bj))
                per_step_launches = int(round(bd["gpu_launches"] / float(bd["steps"])))
            except Exception:
                pass
        n_dev = per_step_launches * 3
        dev = defaultdict(list)
        for r in rows[1:1 + n_dev]:
            v = num(r[vi])
            if v is not None:
                dev[r[ki].split("(")[0]].append(v)
        per_step = {k: sum(v) / 3.0 for k, v in dev.items()}          # time per step (pair_reproj runs 4x per step)
        tot_step = sum(per_step.values())
        f.write("\nDevice-resident 10 M-point steps only (the timed `value` region; first %d launches = 3 steps), "
                "kernel time per step:\n\n| kernel | launches/step | us per step | share of step kernel time |\n|---|---|---|---|\n" % n_dev)
        for k, v in sorted(per_step.items(), key=lambda kv: -kv[1]):
            f.write("| `%s` | %.0f | %.1f | %.3f |\n" % (k.replace("void ", ""), len(dev[k]) / 3.0, v / 1e3, v / tot_step))
    print("wrote", tag + "_launches.md")

# ---- full captures ------------------------------------------------------------------------------------------------
traffic, busy, dur, fp64_exec = {}, {}, {}, {}
with open(os.path.join(dst, tag + "_ncu_full.md"), "w") as f:
    f.write("# `ncu --set full --clock-control none --import-source on` captures (%s)\n\n" % tag)
    f.write("One launch per kernel after warm-up.  Times under the profiler are not bench values.\n")
    reps = set(glob.glob(os.path.join(src, "full_*.ncu-rep"))) | \
        {f[:-len(".raw.csv")] + ".ncu-rep" for f in glob.glob(os.path.join(src, "full_*.raw.csv"))}
    for rep in sorted(reps):
        raw = ncu_csv(rep, "raw")
        if len(raw) < 3:
            continue
        h, u, v = raw[0], raw[1], raw[2]
        col = {n: i for i, n in enumerate(h)}
        name = v[col["Kernel Name"]]
        f.write("\n## %s\n\n`%s`\n\n| metric | value | unit |\n|---|---|---|\n" % (os.path.basename(rep), name))
        for m in METRICS:
            if m in col:
                f.write("| %s | %s | %s |\n" % (m, v[col[m]], u[col[m]]))

        def bytes_of(m):
            x = num(v[col[m]]); unit = u[col[m]]
            return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        grid = num(v[col["launch__grid_size"]])
        tr = bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum")
        f.write("| **DRAM traffic per launch** | %.4g | byte |\n" % tr)
        key = os.path.basename(rep)[len("full_"):-len(".ncu-rep")]
        traffic[key] = tr
        busy[key] = num(v[col["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]])
        dur[key] = num(v[col["gpu__time_duration.sum"]])
        # source page: top lines by sampled stalls
        srcp = ncu_csv(rep, "source")
        while srcp and srcp[0][0] != "Address":
            srcp.pop(0)
        if srcp and "Instructions Executed" in srcp[0]:
            # executed FP64 instructions and how many of them fetch three distinct 64-bit registers (3 issue cycles
            # instead of 2: register-file banking, tools/sass_operands.py)
            hh = srcp[0]
            ci, ce = hh.index("Source"), hh.index("Instructions Executed")
            prev, hist, all_inst = {}, collections.Counter(), 0.0
            for r in srcp[1:]:
                n_exec = num(r[ce]) if len(r) > ce else None
                if not n_exec:
                    continue
                txt = re.sub(r"^@!?U?P\d+\s+", "", r[ci].strip())
                mm = re.match(r"([A-Z0-9_.]+)\s+(.*)", txt)
                if not mm:
                    continue
                all_inst += n_exec
                if mm.group(1).split(".")[0] in ("DFMA", "DMUL", "DADD", "DSETP"):
                    k3, prev = sass_operands.reads(mm.group(1), mm.group(2), prev)
                    hist[k3] += n_exec
                else:
                    prev = {}
            fp = sum(hist.values())
            if fp:
                fp64_exec[key] = {"warp_instr": fp, "all_warp_instr": all_inst, "three_source_share": hist[3] / fp,
                                  "issue_cycles_over_minimum": sum(max(2, k) * c for k, c in hist.items()) / (2.0 * fp)}
                f.write("\nExecuted: %.4g warp instructions, %.4g of them FP64 (%.1f %%); %.1f %% of the FP64 instructions read three "
                        "distinct 64-bit registers (3 issue cycles instead of 2) -> FP64 issue cycles = %.3f x the 2-cycle minimum.\n"
                        % (all_inst, fp, 100 * fp / all_inst, 100 * hist[3] / fp, fp64_exec[key]["issue_cycles_over_minimum"]))
        if srcp:
            hh = srcp[0]
            try:
                si = hh.index("# Samples") if "# Samples" in hh else hh.index("Sampling Data (All)")
            except ValueError:
                si = None
            s_src = hh.index("Source") if "Source" in hh else 1
            if si is not None:
                ranked = sorted((r for r in srcp[1:] if len(r) > si and num(r[si])), key=lambda r: -num(r[si]))[:12]
                total = sum(num(r[si]) or 0 for r in srcp[1:] if len(r) > si)
                f.write("\nTop SASS lines by warp-state samples (total %d):\n\n| samples | share | instruction |\n|---|---|---|\n" % total)
                for r in ranked:
                    f.write("| %s | %.3f | `%s` |\n" % (r[si], num(r[si]) / max(total, 1), r[s_src].strip()[:110]))
print("wrote", tag + "_ncu_full.md")

for fn in glob.glob(os.path.join(src, "bench_*.json")) + glob.glob(os.path.join(src, "sweep_*.jsonl")) + \
        glob.glob(os.path.join(src, "*.log")):
    if os.path.getsize(fn) and "full_" not in os.path.basename(fn) and "launches_bench" not in fn:
        shutil.copy(fn, os.path.join(dst, tag + "_" + os.path.basename(fn)))
print(json.dumps(traffic))
# profiles/traffic.json: DRAM bytes per point of the kernels bench.py reports (10 M-point captures of this round)
names = {"ls": "linear_LS", "iter_eval": "iterative_LS", "eigen_eval": "linear_eigen", "poly_eval": "polynomial"}
if all(k in traffic for k in names):
    tj = {"source": "profiles/%s_ncu_full.md (ncu --set full, 10 M points, dram__bytes_read.sum + dram__bytes_write.sum per launch "
                    "/ 1e7 points; linear_LS: the plain HBM-bound kernel of the roofline entry; the three FP64-bound kernels "
                    "with the evaluation epilogue and the good mask)" % tag}
    for k, nme in names.items():
        tj[nme] = traffic[k] / 1e7
    tj["linear_LS_eval"] = traffic.get("ls_eval", 0) / 1e7
    tj["fp64_pipe_busy"] = {"source": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active, profiles/%s_ncu_full.md" % tag}
    for k, nme in names.items():
        tj["fp64_pipe_busy"][nme] = round(busy[k] / 100.0, 3)
    if all(k in fp64_exec for k in names):
        tj["fp64_executed"] = {"source": "executed-instruction counts of the source page of the same captures (10 M points): FP64 "
                                         "instructions per point, and the share of them with three distinct register sources"}
        for k, nme in list(names.items()) + [("ls_eval", "linear_LS_eval")]:
            if k in fp64_exec:
                e = fp64_exec[k]
                tj["fp64_executed"][nme] = {"fp64_instr_per_point": round(e["warp_instr"] * 32 / 1e7, 1),
                                            "instr_per_point": round(e["all_warp_instr"] * 32 / 1e7, 1),
                                            "three_source_share": round(e["three_source_share"], 3),
                                            "issue_cycles_over_minimum": round(e["issue_cycles_over_minimum"], 3)}
    if "ls_100M" in traffic:
        tj["linear_LS_100M_bytes_per_point"] = traffic["ls_100M"] / 1e8
    with open(os.path.join(dst, "traffic.json"), "w") as f:
        json.dump(tj, f, indent=1)
    print("wrote traffic.json")
