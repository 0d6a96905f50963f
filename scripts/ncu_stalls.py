"""Warp-stall breakdown + hottest source lines of one kernel from an .ncu-rep (read here, no GPU needed):
   python scripts/ncu_stalls.py file.ncu-rep [top_lines]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur = None; agg = []; stalls = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 3 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-": continue
    try:
        ie = int(r[hdr.index("Instructions Executed")]); te = int(r[hdr.index("Thread Instructions Executed")]); sm = int(r[hdr.index("# Samples")])
    except ValueError: continue
    agg.append((sm, ie, te, cur, r[0], r[1].strip()[:100]))
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            try: stalls[h] = stalls.get(h, 0) + int(r[i])
            except ValueError: pass
ti = sum(a[1] for a in agg); ts = sum(a[0] for a in agg)
print(f"warp instructions {ti}  samples {ts}")
print("stalls: " + "  ".join(f"{k[6:]} {v / max(ts, 1) * 100:.1f}%" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]))
for a in sorted(agg, reverse=True)[:top]:
    print(f"{a[0] / ts * 100:5.1f}% smp {a[1] / ti * 100:5.1f}% inst  lanes {a[2] / max(a[1], 1):5.1f}  {a[3]}:{a[4]}  {a[5]}")
