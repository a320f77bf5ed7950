#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel medians and shares.
    python tools/launch_summary.py gpurun_out/launches.csv profiles/rNN_launches.csv "header note"
Share = median / sum of medians of the kernels that make up one training step (eval-only instantiations excluded)."""
import csv, statistics, sys, re
from collections import defaultdict
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
iK, iV = hdr.index("Kernel Name"), hdr.index("Metric Value")
t = defaultdict(list)
for r in rd:
    name = re.sub(r"\(.*", "", r[iK]).replace("void ", "").replace("vqb200::", "")
    if "elementwise" in name or "at::" in name or "distribution" in name:
        continue
    t[name].append(float(r[iV]) / 1e3)
out = open(sys.argv[2], "w")
out.write(f"# {sys.argv[3] if len(sys.argv) > 3 else ''}\n")
out.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: compare shares, not absolutes)\n")
med = {k: statistics.median(v) for k, v in t.items()}
step = {k: m for k, m in med.items() if "<0, 0," not in k and "<false, false" not in k and "lookup" not in k}
tot = sum(step.values())
out.write("kernel,launches,median_us,min_us,share_of_train_step\n")
for k in sorted(med, key=lambda k: -med[k]):
    out.write(f"{k},{len(t[k])},{med[k]:.1f},{min(t[k]):.1f},{(med[k] / tot if k in step else float('nan')):.3f}\n")
out.close()
print(open(sys.argv[2]).read())
