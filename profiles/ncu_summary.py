"""Condense an `ncu -i X.ncu-rep --page raw --csv` export to the metrics the roofline discussion uses.
usage: python profiles/ncu_summary.py raw.csv "note" > profiles/<name>.csv"""
import csv
import sys

KEEP = ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit",
        "launch__shared_mem_per_block_dynamic", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_fp64", "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_lsu")
rows = list(csv.reader(open(sys.argv[1])))
h, u = rows[0], rows[1]
print("#", sys.argv[2] if len(sys.argv) > 2 else "")
print("kernel,metric,unit,value")
for v in rows[2:]:
    name = v[h.index("Kernel Name")].split("(")[0]
    for i, m in enumerate(h):
        if any(m.startswith(k) for k in KEEP) and v[i] not in ("", "0", "0.000000"):
            print(f"{name},{m},{u[i]},{v[i]}")
