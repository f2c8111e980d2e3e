#!/usr/bin/env python
"""Turn the raw ncu output of one gpurun call (gpurun_out/) into the committed summaries.
usage: python profiles/refresh.py RAW_TAG OUT_TAG      e.g.  refresh.py r01j r01b
  gpurun_out/launches_<RAW_TAG>.csv             -> profiles/<OUT_TAG>_launches.csv, <OUT_TAG>_launch_shares.csv
  gpurun_out/prof_<RAW_TAG>_<kernel>.ncu-rep    -> profiles/<OUT_TAG>_<kernel>_full.txt (summarize.py selection)"""
import csv, glob, io, os, re, subprocess, sys
raw, out = sys.argv[1], sys.argv[2]
here = os.path.dirname(os.path.abspath(__file__)); root = os.path.dirname(here)
src = os.path.join(root, "gpurun_out", f"launches_{raw}.csv")
if os.path.exists(src):
    lines = [l for l in open(src) if l.startswith('"')]
    open(os.path.join(here, f"{out}_launches.csv"), "w").writelines(lines)
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = {}
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r["Metric Value"]) / 1e6
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(here, f"{out}_launch_shares.csv"), "w") as f:
        f.write(f"# ncu launch list summary ({len(rows)} launches, {tot:.2f} ms of profiled device time; per-launch times are cold-cache, serialised)\n")
        f.write("kernel,launches,total_ms,share\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{v[0]},{v[1]:.3f},{v[1] / tot:.3f}\n")
for rep in sorted(glob.glob(os.path.join(root, "gpurun_out", f"prof_{raw}_*.ncu-rep"))):
    k = os.path.basename(rep)[len(f"prof_{raw}_"):-len(".ncu-rep")]
    txt = subprocess.run([sys.executable, os.path.join(here, "summarize.py"), rep], capture_output=True, text=True).stdout
    open(os.path.join(here, f"{out}_{k}_full.txt"), "w").write(txt)
    print("wrote", f"{out}_{k}_full.txt")
