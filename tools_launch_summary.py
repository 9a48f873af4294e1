#!/usr/bin/env python
"""Per-kernel summary of ONE training step out of an `ncu --metrics gpu__time_duration.sum --csv` launch list
(a step = the launches from one stem im2col kernel to the next).   python tools_launch_summary.py list.csv [out.md] [title]"""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"]) * (1e-3 if r["Metric Unit"] == "ns" else 1.0)))     # us
marks = [i for i, (k, _t) in enumerate(rows) if "im2col_7x7s2" in k]
if len(marks) < 2:
    raise SystemExit("the list does not hold one whole step (need two stem im2col launches)")
step = rows[marks[-2]:marks[-1]]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("hgb::", "").replace("(int)", "").replace("(bool)", "")
    if name.startswith("at::") or "at::native" in name or "elementwise" in name:
        return "torch (fill / copy / reduce helpers)"
    return name[:110]


agg = OrderedDict()
for k, t in step:
    a = agg.setdefault(short(k), [0, 0.0, 1e30])
    a[0] += 1
    a[1] += t
    a[2] = min(a[2], t)
tot = sum(a[1] for a in agg.values())
ours = sum(a[1] for k, a in agg.items() if not k.startswith("torch") and "nccl" not in k.lower())
out = [f"# {sys.argv[3] if len(sys.argv) > 3 else 'ncu launch list of ONE training step'}", "",
       f"Source list: `{sys.argv[1].split('/')[-1]}`; one step = launches {marks[-2]}..{marks[-1] - 1} of the list (stem im2col to stem im2col). "
       "Under ncu every kernel runs alone (no lanes, no programmatic-dependent-launch overlap), cold caches: compare SHARES, not absolute times.", "",
       f"Sum of kernel durations: {tot / 1e3:.2f} ms over {len(step)} launches; kernels of libhgb200.so: {100 * ours / tot:.1f} % of that time.", "",
       "| kernel | launches | total ms | share | avg us | min us |", "|---|---:|---:|---:|---:|---:|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {a[0]} | {a[1] / 1e3:.3f} | {100 * a[1] / tot:.1f}% | {a[1] / a[0]:.1f} | {a[2]:.1f} |")
text = "\n".join(out) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
else:
    print(text)
