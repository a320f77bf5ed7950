#!/usr/bin/env python
"""Per-source-line view of an `ncu --set full --import-source on` report: warp-stall samples, executed instructions and
shared-memory wavefronts per CUDA source line (all inlined copies summed), plus the LSU totals of every launch.

    python tools/ncu_lines.py report.ncu-rep [tiles] [top]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
tiles = float(sys.argv[2]) if len(sys.argv) > 2 else 8192.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 28


def page(p, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", p, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


raw = page("raw")
hdr = raw[0]
WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread"]
for r in raw[2:]:
    print("==", r[hdr.index("Kernel Name")][:70])
    for k in WANT:
        if k in hdr:
            v = r[hdr.index(k)]
            try:
                fv = float(v.replace(",", ""))
                extra = "   (%.1f / tile)" % (fv / tiles) if ("wavefronts" in k or "conflicts" in k or "inst_executed.sum" in k) and "pct" not in k else ""
            except ValueError:
                extra = ""
            print("   %-78s %s%s" % (k, v, extra))

rows = page("source", ["--print-source", "sass,cuda"])
blocks = [i for i, r in enumerate(rows) if r and r[0] == "Function Name"]
seen = 0
agg_by_kernel = {}
for n, bi in enumerate(blocks):
    name = rows[bi][1]
    h = rows[bi + 1]
    end = blocks[n + 1] - 1 if n + 1 < len(blocks) else len(rows)
    iw, ie, isamp = h.index("L1 Wavefronts Shared"), h.index("Instructions Executed"), h.index("# Samples")
    agg = agg_by_kernel.setdefault(name, {})
    cur = None
    for r in rows[bi + 2:end]:
        if len(r) <= iw:
            continue
        if r[0] != "":
            cur = (rows[bi - 1][1].split("/")[-1] if rows[bi - 1] and rows[bi - 1][0] == "File Path" else "", int(r[0]), r[1].strip()[:95])
            agg.setdefault(cur, [0, 0, 0])
            continue
        if r[2] in ("...", "") or cur is None:
            continue
        try:
            a = agg[cur]
            a[0] += int(r[iw] or 0); a[1] += int(r[ie] or 0); a[2] += int(r[isamp] or 0)
        except ValueError:
            pass
for name, agg in agg_by_kernel.items():
    tots = sum(a[2] for a in agg.values()) or 1
    print("\n==== %s\n  shared wavefronts / tile %.0f, instructions / tile %.0f, samples %d" %
          (name[:90], sum(a[0] for a in agg.values()) / tiles, sum(a[1] for a in agg.values()) / tiles, tots))
    print("  -- by stall samples")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
        print("   %-14s L%-5d %5.1f%%  inst/tile %7.1f  wf/tile %7.1f | %s" % (k[0][:14], k[1], 100.0 * a[2] / tots, a[1] / tiles, a[0] / tiles, k[2]))
    print("  -- by shared-memory wavefronts")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:12]:
        print("   %-14s L%-5d wf/tile %7.1f  inst/tile %7.1f | %s" % (k[0][:14], k[1], a[0] / tiles, a[1] / tiles, k[2]))
