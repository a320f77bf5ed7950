#!/usr/bin/env python
"""Stall-reason breakdown of a source-line range of an ncu report (--import-source on).
   python tools/ncu_stalls.py report.ncu-rep first_line last_line"""
import csv, io, subprocess, sys
rep, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks = [i for i, r in enumerate(rows) if r and r[0] == "Function Name"]
tot = {}
lines = {}
for n, bi in enumerate(blocks):
    h = rows[bi + 1]
    end = blocks[n + 1] - 1 if n + 1 < len(blocks) else len(rows)
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    isamp, ie = h.index("# Samples"), h.index("Instructions Executed")
    fname = rows[bi - 1][1] if rows[bi - 1] and rows[bi - 1][0] == "File Path" else ""
    if not fname.endswith("vq_assign_tc.cu"):
        continue
    cur = None
    for r in rows[bi + 2:end]:
        if len(r) <= isamp:
            continue
        if r[0] != "":
            cur = int(r[0]); continue
        if r[2] in ("...", "") or cur is None or not (lo <= cur <= hi):
            continue
        try:
            L = lines.setdefault(cur, [0, 0, {}])
            L[0] += int(r[isamp] or 0); L[1] += int(r[ie] or 0)
            for i, c in stall_cols:
                v = int(r[i] or 0)
                if v:
                    L[2][c] = L[2].get(c, 0) + v
                    tot[c] = tot.get(c, 0) + v
        except ValueError:
            pass
S = sum(tot.values()) or 1
print("lines %d..%d: %d samples, %d instructions" % (lo, hi, sum(l[0] for l in lines.values()), sum(l[1] for l in lines.values())))
for c, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print("   %-28s %6d  %5.1f%%" % (c, v, 100.0 * v / S))
print("top lines:")
for ln, L in sorted(lines.items(), key=lambda kv: -kv[1][0])[:14]:
    top = ", ".join("%s %d" % (c.replace("stall_", ""), v) for c, v in sorted(L[2].items(), key=lambda kv: -kv[1])[:3])
    print("   L%-5d samples %5d inst %8d | %s" % (ln, L[0], L[1], top))
