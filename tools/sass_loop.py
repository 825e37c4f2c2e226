#!/usr/bin/env python
"""Static SASS census of a kernel's hottest loop (the backward branch spanning the most FFMAs): instruction count
and opcode histogram.   python tools/sass_loop.py <kernel-name-substring> [lib.so]"""
import collections
import re
import subprocess
import sys

name = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else "md_rdm_b200/librdm_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
body = [f for f in funcs if name in f.split("\n")[0]]
assert body, "kernel not found"
ins = []
for line in body[0].split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
best = None
cands = []
for i, (addr, s) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", s)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < addr:
            j = next(k for k, (a, _) in enumerate(ins) if a >= tgt)
            n_ffma = sum(1 for _, t in ins[j:i + 1] if "FFMA" in t)
            if n_ffma >= 150:
                cands.append((n_ffma, j, i))
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cands.sort(key=lambda c: c[2] - c[1])
print("FMA-heavy loops (instructions):", [c[2] - c[1] + 1 for c in cands])
best = cands[which]
n_ffma, j, i = best
loop = ins[j:i + 1]
c = collections.Counter()
for _, s in loop:
    t = s.split()
    op = t[1] if t[0].startswith("@") else t[0]
    c[op.split(".")[0]] += 1
print(f"{name}: loop of {len(loop)} instructions ({hex(loop[0][0])}..{hex(loop[-1][0])})")
print(sorted(c.items(), key=lambda kv: -kv[1]))
