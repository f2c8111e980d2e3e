#!/usr/bin/env python
"""Attribute executed warp-instructions of one kernel to CUDA source lines by joining the ncu SASS
page (per-instruction counts, in program order) with `nvdisasm -g` line info of the same cubin.
usage: python profiles/by_line.py rep cubin kernel_substring units_per_launch [min_per_unit]"""
import csv, io, re, subprocess, sys
rep, cubin, kname, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
thr = float(sys.argv[5]) if len(sys.argv) > 5 else 20
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr = rows[1]
iS, iI = hdr.index("Source"), hdr.index(sys.argv[6] if len(sys.argv) > 6 else "Instructions Executed")
counts = [(r[iS].strip(), float(r[iI])) for r in rows[2:]]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# locate the kernel's section
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kname in l)
cur = None; ins = []
for l in dis[start + 1:]:
    if l.startswith("//---------------------"): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        # inlined-from info is on the same comment; keep innermost
        continue
    m2 = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m2: ins.append((cur, m2.group(1)))
if len(ins) != len(counts):
    print(f"warning: {len(ins)} disassembled instructions vs {len(counts)} profiled", file=sys.stderr)
agg = {}
for (loc, _), (_, c) in zip(ins, counts):
    agg[loc] = agg.get(loc, 0.0) + c
tot = sum(c for _, c in counts)
src_cache = {}
def src(loc):
    if not loc: return ""
    f, n = loc
    if f not in src_cache:
        import glob
        cands = glob.glob(f"/root/repo/savont_b200/csrc/{f}") + glob.glob(f"/root/repo/**/{f}", recursive=True)
        src_cache[f] = open(cands[0]).read().splitlines() if cands else []
    L = src_cache[f]
    return L[n - 1].strip()[:100] if 0 < n <= len(L) else ""
print(f"total {tot / units:.1f} instr/unit")
for loc, c in sorted(agg.items(), key=lambda kv: (kv[0] or ("", 0))):
    if c / units >= thr:
        print(f"{c / units:8.1f}  {loc[0] if loc else '?'}:{loc[1] if loc else 0:4d}  {src(loc)}")
