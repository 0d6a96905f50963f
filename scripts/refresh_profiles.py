"""Turns gpurun_out/<tag>_* captures (scripts/capture_profiles.sh) into the tracked files under profiles/.
   python scripts/refresh_profiles.py r01"""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

# launch list -> per-kernel totals and shares
rows = [r for r in csv.reader(open(f"{G}/{tag}_launches.csv", errors="replace")) if len(r) > 10]
hdr = rows[0]; kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value"); mu = hdr.index("Metric Unit")
tot = {}
for r in rows[1:]:
    v = float(r[mv].replace(",", "")); v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[mu], 1e-6)
    n = r[kn].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    t = tot.setdefault(n, [0, 0.0]); t[0] += 1; t[1] += v
allms = sum(t[1] for t in tot.values())
with open(f"{P}/{tag}_launches_summary.txt", "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none, python scripts/profile_frame.py 2 (scene load + BLAS/TLAS build + 2 frames of the bench workload)\n")
    for n, t in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{n:60s} n={t[0]:5d} total={t[1]:9.3f} ms share={t[1] / allms:.3f}\n")
shutil.copy(f"{G}/{tag}_launches.csv", f"{P}/{tag}_launches.csv")
shutil.copy(f"{G}/{tag}_bench_n1.json", f"{P}/{tag}_bench_n1.json")

traffic = {}
for k, out in (("k_trace", "trace"), ("k_shade", "shade"), ("misc", "misc")):
    rep = f"{G}/{tag}_prof_{k}.ncu-rep"
    if not os.path.exists(rep): continue
    txt = subprocess.run([sys.executable, f"{ROOT}/scripts/ncu_summary.py", rep], capture_output=True, text=True).stdout
    open(f"{P}/{tag}_ncu_{out}.txt", "w").write(txt)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines())); h, u = rr[0], rr[1]
    for r in rr[2:]:
        name = r[h.index("Kernel Name")].split("(")[0].split("<")[0].replace("void ", "").split("::")[-1]
        def val(m):
            i = h.index(m); x = float(r[i].replace(",", ""))
            return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[i]]
        traffic.setdefault(name, []).append(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
tj = {k: sum(v) / len(v) for k, v in traffic.items()}
tj["note"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu). k_trace / k_shade: mean over ALL launches of one whole frame (the launches roofline.achieved averages over); the others: mean over the captured launches of a --set full pass"
# whole-frame traffic of the wavefront kernels: every launch of one frame, metrics-only pass
ft = f"{G}/{tag}_frame_traffic.csv"
if os.path.exists(ft):
    rows = [r for r in csv.reader(open(ft, errors="replace")) if len(r) > 10]
    h = rows[0]; kn, mn, mv, mu, idc = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
    per = {}
    for r in rows[1:]:
        if not r[mn].startswith("dram__bytes"): continue
        name = r[kn].split("(")[0].split("<")[0].replace("void ", "").split("::")[-1]
        x = float(r[mv].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[mu]]
        per.setdefault(name, {}).setdefault(r[idc], 0.0)
        per[name][r[idc]] += x
    for name, d in per.items():
        tj[name] = sum(d.values()) / len(d)
        tj[name + "_launches_in_frame"] = len(d)
        tj[name + "_frame_total_bytes"] = sum(d.values())
    shutil.copy(ft, f"{P}/{tag}_frame_traffic.csv")
json.dump(tj, open(f"{P}/traffic.json", "w"), indent=1)
print(open(f"{P}/{tag}_launches_summary.txt").read()); print(json.dumps(tj, indent=1))

subprocess.run([sys.executable, f"{ROOT}/scripts/ncu_table.py", tag])
