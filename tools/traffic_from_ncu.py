"""Fill profiles/traffic.json from ncu launch lists (csv, --metrics dram__bytes_read.sum,dram__bytes_write.sum,...):
    python tools/traffic_from_ncu.py <key_prefix> <kind> <csv> <source label>
Kernel names containing '<..., 1, ...>' in the STATS slot are the training launches, the others eval."""
import collections
import csv
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prefix, kind, path, source = sys.argv[1:5]
rows = list(csv.reader(open(path)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ix = {n: i for i, n in enumerate(H)}
agg = collections.defaultdict(dict)
for r in rows[hdr + 1:]:
    if len(r) < len(H):
        continue
    agg[(r[ix["ID"]], r[ix["Kernel Name"]])][r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
by_mode = collections.defaultdict(list)
for (_, name), m in agg.items():
    args = name[name.index("<") + 1:name.index(">")].split(",")
    stats = args[1].strip() in ("1", "true", "(bool)1")
    by_mode["train" if stats else "eval"].append((name, m))
tj_path = os.path.join(ROOT, "profiles", "traffic.json")
tj = json.load(open(tj_path))
for mode, lst in by_mode.items():
    rd = statistics.median(m["dram__bytes_read.sum"] for _, m in lst)
    wr = statistics.median(m["dram__bytes_write.sum"] for _, m in lst)
    us = statistics.median(m["gpu__time_duration.sum"] for _, m in lst) / 1e3
    tj[f"{prefix}_{mode}_{kind}"] = {"kernel": lst[0][0][:60], "bytes": int(rd + wr), "read": int(rd), "write": int(wr),
                                     "launches": len(lst), "median_us_under_ncu": round(us, 1), "source": source}
    print(f"{prefix}_{mode}_{kind}: read {rd / 1e6:.1f} MB write {wr / 1e6:.1f} MB, {us:.1f} us under ncu, {len(lst)} launches")
json.dump(tj, open(tj_path, "w"), indent=2)
