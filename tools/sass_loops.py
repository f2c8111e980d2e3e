#!/usr/bin/env python
"""List the backward-branch loops of one kernel in a built .so with their opcode histograms (pipe split: ALU vs FMA).
usage: python tools/sass_loops.py lib.so kernel_substring [min_instr]"""
import re, subprocess, sys, collections
so, pat = sys.argv[1], sys.argv[2]
mn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur = None; ins = {}
for l in txt.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m: cur = m.group(1); ins[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and cur: ins[cur].append((int(m.group(1), 16), m.group(2).strip()))
FMA = ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "IDP")
ALU = ("LOP3", "SHF", "IADD3", "VIADD", "ISETP", "SEL", "PRMT", "VIMNMX", "LEA", "POPC", "FLO", "SGXT", "BMSK", "IABS", "PLOP3", "FSETP", "FSEL", "MOV", "I2I", "IADD")
for k, v in ins.items():
    if pat not in k: continue
    print("==", k, len(v), "instructions")
    addr = {a: i for i, (a, _) in enumerate(v)}
    for i, (a, t) in enumerate(v):
        m = re.search(r"BRA(?:\.\w+)*\s+(?:\w+,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) in addr and addr[int(m.group(1), 16)] <= i:
            j = addr[int(m.group(1), 16)]
            if i - j + 1 < mn: continue
            ops = collections.Counter()
            for _, tt in v[j:i + 1]:
                tt = re.sub(r"^@!?U?P\d+\s+", "", tt)
                ops[tt.split()[0].split(".")[0]] += 1
            alu = sum(c for o, c in ops.items() if o in ALU); fma = sum(c for o, c in ops.items() if o in FMA)
            print(f"  loop [{j}..{i}] {i - j + 1} instr: ALU {alu} FMA {fma} other {i - j + 1 - alu - fma}  ", dict(ops.most_common()))
