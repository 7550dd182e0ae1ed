"""Brief view of an exported `ncu --page raw --csv` file: the metrics that decide what bounds a kernel.
   python tools/ncu_brief.py gpurun_out/<tag>/full_<kernel>.raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
units = rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "")[:80], "grid", d.get("launch__grid_size"), "regs", d.get("launch__registers_per_thread"))
    keys = ["gpu__time_duration.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg",
            "smsp__cycles_active.avg", "sm__cycles_active.avg"]
    keys += [k for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
    keys += [k for k in hdr if k.startswith("smsp__inst_executed_pipe_") and k.endswith(".sum")]
    for k in keys:
        if k in d and d[k] not in ("", "0", "0.00"):
            print("   %-86s %s %s" % (k, d[k], units[hdr.index(k)]))
