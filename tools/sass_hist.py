"""Opcode histogram + hottest SASS lines of an ncu report's source page.
  python tools/sass_hist.py report.ncu-rep [kernel-index]"""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
# the output holds one table per kernel, each preceded by a "Kernel Name",... line
blocks, cur = [], None
for ln in txt.splitlines():
    if ln.startswith('"Kernel Name"'):
        cur = [ln]; blocks.append(cur)
    elif cur is not None:
        cur.append(ln)
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
b = blocks[which]
print(b[0][:200])
rows = list(csv.DictReader(io.StringIO("\n".join(b[1:]))))
c, st, tot, stot = Counter(), Counter(), 0, 0
lines = []
for r in rows:
    src = r["Source"].strip()
    n = int(float(r["Instructions Executed"] or 0))
    s = int(float(r["Warp Stall Sampling (All Samples)"] or 0))
    toks = src.split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0] if toks else "?"
    c[op] += n; tot += n; st[op] += s; stot += s
    lines.append((s, n, src))
print("total warp instructions", tot, "stall samples", stot)
for k, v in c.most_common(22):
    print(f"{k:10s} {v:12d} {100*v/tot:5.1f}%   stall {100*st[k]/max(stot,1):5.1f}%")
print("--- hottest lines by stall samples")
for s, n, src in sorted(lines, reverse=True)[:25]:
    print(f"{100*s/max(stot,1):5.1f}% {n:10d}  {src[:110]}")
