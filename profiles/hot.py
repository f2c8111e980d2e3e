#!/usr/bin/env python
"""Per-SASS-line instruction counts and stall samples from an .ncu-rep (source page), grouped
into contiguous hot regions.  usage: python profiles/hot.py rep [top_n]"""
import csv, io, subprocess, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
iS, iN, iI = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
body = rows[2:]
def f(x):
    try: return float(x)
    except Exception: return 0.0
totI = sum(f(r[iI]) for r in body); totS = sum(f(r[iN]) for r in body)
print(f"total warp-instructions {totI:.3e}, samples {totS:.0f}, sass lines {len(body)}")
idx = sorted(range(len(body)), key=lambda i: -f(body[i][iI]))[:topn]
for i in sorted(idx):
    r = body[i]
    print(f"{i:5d} {100*f(r[iI])/totI:5.2f}%I {100*f(r[iN])/max(totS,1):5.2f}%S  {r[iS][:100]}")
