"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X ...`)
into a per-kernel table for ONE step (the launches between the last two stem_im2col_kernel launches).
  python tools/summarize_launches.py gpurun_out/launches.csv [marker-kernel-substring]"""
import csv
import sys
from collections import OrderedDict

path = sys.argv[1]
marker = sys.argv[2] if len(sys.argv) > 2 else "stem_s2d"
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", "")) / 1e6))  # ns -> ms
marks = [i for i, (k, _) in enumerate(rows) if marker in k]
assert len(marks) >= 2, f"need two '{marker}' launches, found {len(marks)} in {len(rows)} rows"
step = rows[marks[-2]:marks[-1]]
agg = OrderedDict()
for k, ms in step:
    name = k.split("(")[0].replace("void ", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(ms for _, ms in step)
print(f"One step = launches between the last two `{marker}` launches: {len(step)} launches, {tot:.2f} ms summed.\n")
print("| kernel | launches | ms | share |\n|---|---|---|---|")
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name[-70:]}` | {n} | {ms:.3f} | {100 * ms / tot:.1f}% |")
