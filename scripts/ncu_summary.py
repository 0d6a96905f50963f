"""Summarises an .ncu-rep (read here, no GPU needed): python scripts/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "smsp__sass_average_branch_targets_threads_uniform.pct", "sass__inst_executed_local_loads",
        "sass__inst_executed_local_stores", "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
name_col = hdr.index("Kernel Name")
for r in data:
    print("==", r[name_col].split("(")[0])
    for i, h in enumerate(hdr):
        if h in WANT:
            print(f"   {h:70s} {r[i]:>18s} {units[i]}")
