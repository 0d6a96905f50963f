"""Per-kernel ncu table (HBM and L2 GB/s against the measured peaks, SM issue utilisation, warp divergence) from the
gpurun_out/<tag>_prof_*.ncu-rep captures:  python scripts/ncu_table.py r01  ->  profiles/<tag>_ncu_table.md"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
def find(d, keys):
    if isinstance(d, dict):
        for k, v in d.items():
            if any(x in k.lower() for x in keys) and isinstance(v, (int, float)): return float(v)
            r = find(v, keys)
            if r: return r
    return None
hbm_peak = find(peaks, ("hbm", "copy", "dram")) or 6548.8
rows = {}
M = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
     "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
SC = {"sector": 1, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1}
for k in ("k_trace", "k_shade", "misc"):
    rep = os.path.join(ROOT, "gpurun_out", f"{tag}_prof_{k}.ncu-rep")
    if not os.path.exists(rep): continue
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(out.splitlines())); h, u = rr[0], rr[1]
    for r in rr[2:]:
        name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "").split("::")[-1]
        v = {}
        for m in M:
            if m in h:
                i = h.index(m); v[m] = float(r[i].replace(",", "")) * SC.get(u[i], 1)
        rows.setdefault(name, []).append(v)
lines = [f"# {tag} — ncu --set full --clock-control none, per launch (mean of the captured launches; k_trace / k_shade: waves 10-11 of a bench frame, ~2.07 M rays each)",
         "", f"HBM peak used: {hbm_peak:.1f} GB/s (MEASURED_PEAKS.json). ncu replays every kernel cold and serialised, so durations are longer than in-frame (bench.py reports the in-frame ones);",
         "the ratios are what this table is for. L2 GB/s = lts__t_sectors x 32 B / duration (L2 % = lts__throughput of peak). Divergence = active threads per executed warp instruction (32 = none).", "",
         "| kernel | launches | duration µs | HBM GB/s | % of HBM peak | L2 GB/s | L2 % of peak | SM issue active % | threads / warp instr | warps active % | L1 hit % | L2 hit % | regs | warp instr (M) |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
for name, vs in rows.items():
    mean = {m: sum(x.get(m, 0) for x in vs) / len(vs) for m in M}
    t = mean["gpu__time_duration.sum"]
    hbm = (mean["dram__bytes_read.sum"] + mean["dram__bytes_write.sum"]) / t / 1e9
    l2 = mean["lts__t_sectors.sum"] * 32 / t / 1e9
    lines.append(f"| `{name}` | {len(vs)} | {t * 1e6:.1f} | {hbm:.0f} | {100 * hbm / hbm_peak:.1f} | {l2:.0f} | {mean['lts__throughput.avg.pct_of_peak_sustained_elapsed']:.1f} | {mean['smsp__issue_active.avg.pct_of_peak_sustained_active']:.1f} | "
                 f"{mean['smsp__thread_inst_executed_per_inst_executed.ratio']:.1f} | {mean['sm__warps_active.avg.pct_of_peak_sustained_active']:.1f} | "
                 f"{mean['l1tex__t_sector_hit_rate.pct']:.1f} | {mean['lts__t_sector_hit_rate.pct']:.1f} | {mean['launch__registers_per_thread']:.0f} | {mean['smsp__inst_executed.sum'] / 1e6:.1f} |")
open(os.path.join(ROOT, "profiles", f"{tag}_ncu_table.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
