#!/usr/bin/env python
"""Summarise an `ncu --set full` report (.ncu-rep) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_name.txt ["free-form note"]

Reads the report with `ncu -i ... --page raw --csv` / `--page source --csv` (works without a GPU) and keeps
the counters the roofline argument needs: duration, DRAM bytes / throughput, tensor-pipe activity, issue
utilisation, registers, plus the opcode mix and the top stall reasons of the SASS.
"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "l1tex__data_bank_conflicts_pipe_lsu.sum",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu summary of {rep}", f"# {note}", ""]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines.append(f"== kernel: {d.get('Kernel Name', '?')}   (launch id {d.get('ID', '?')})")
        for k in KEEP:
            for h in hdr:
                if h == k or h.endswith("." + k):
                    lines.append(f"  {k:72s} {d[h]:>18s} {u[h]}")
                    break
        rd, wr, t = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum"), d.get("gpu__time_duration.sum")
        lines.append("")
    # SASS opcode mix + stall reasons (first kernel in the report that has a source page)
    src = ncu_csv(rep, "source")
    if len(src) > 2:
        h = src[1]
        ix = {k: i for i, k in enumerate(h)}
        if "Source" in ix:
            ops, tot = Counter(), 0
            stalls = Counter()
            stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
            for r in src[2:]:
                if len(r) < len(h):
                    continue
                parts = r[ix["Source"]].split()
                if not parts:
                    continue
                op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
                try:
                    n = int(r[ix["Instructions Executed"]] or 0)
                except ValueError:          # the header row of the next profiled launch: only the first one is summarised
                    break
                ops[op.split(".")[0]] += n
                tot += n
                for c in stall_cols:
                    stalls[c] += int(r[ix[c]] or 0)
            lines.append("== SASS opcode mix (warp instructions executed, all passes of the first profiled launch)")
            for op, n in ops.most_common(24):
                lines.append(f"  {op:14s} {n:14d} {100.0 * n / max(tot, 1):6.2f} %")
            lines.append("== warp stall samples")
            st = sum(stalls.values())
            for c, n in stalls.most_common(10):
                lines.append(f"  {c:28s} {n:10d} {100.0 * n / max(st, 1):6.2f} %")
            marks = [m for m in ("UTCHMMA", "UTCQMMA", "UTCMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA")
                     if any(k.startswith(m) for k in ops)]
            lines.append("== Blackwell SASS markers present: " + ", ".join(marks))
    with open(dst, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
