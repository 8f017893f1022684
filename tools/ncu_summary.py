"""Condenses `ncu -i X.ncu-rep --page raw --csv` dumps into the JSON summary kept under profiles/.
Usage: ncu_summary.py out.json raw1.csv [raw2.csv ...]   (one kernel launch per csv row)"""
import csv, json, sys
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__inst_executed.avg.per_cycle_active", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
        "sm__cycles_elapsed.avg.per_second", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
out = []
for path in sys.argv[2:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {"Kernel Name": r[hdr.index("Kernel Name")], "source_csv": path.split("/")[-1]}
        for k in KEYS:
            if k in hdr:
                d[k] = (r[hdr.index(k)] + " " + units[hdr.index(k)]).strip()
        st = {}
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                st[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(float(r[i]), 3)
        d["stalls_per_issue"] = st
        out.append(d)
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(f"{len(out)} launches -> {sys.argv[1]}")
