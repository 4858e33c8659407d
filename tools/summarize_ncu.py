#!/usr/bin/env python3
"""Turns an .ncu-rep (from gpurun_out/) into the text summary committed under profiles/:
key raw metrics per captured launch + the hottest CUDA source lines with lanes active per instruction.
usage: summarize_ncu.py <report.ncu-rep> <out.txt> [note]"""
import collections
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
lines = ["ncu summary of %s" % rep.split("/")[-1], note, ""]
for r in rows[2:]:
    lines.append("kernel: " + (r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"))
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            lines.append("  %-62s %14s %s" % (w, r[i], units[i]))
    lines.append("")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE,
                     text=True).stdout
cur, agg, h = None, collections.OrderedDict(), None
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Line No":
        h = r
        continue
    if h is None or len(r) < 10:
        continue
    if r[2] == "-" and r[0].isdigit():
        try:
            agg[(cur, int(r[0]))] = (int(r[7]), int(r[8]), int(r[6]), r[1].strip()[:88])
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values()) or 1
lines.append("hottest source lines (share of warp instructions executed, lanes active per instruction, stall samples):")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:32]:
    lines.append("  %-16s %4d  %5.1f%%  lanes %5.1f  samples %6d | %s" % (k[0], k[1], 100.0 * v[0] / tot, v[1] / max(v[0], 1), v[2], v[3]))
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
