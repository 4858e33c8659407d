#!/usr/bin/env python3
"""Reads the raw page of .ncu-rep captures (one launch each) and writes the figures bench.py's roofline block is
computed from — warp instructions per ray, lanes active per issued instruction, issue-active %, DRAM bytes — into
profiles/r02_ncu_metrics.json.  usage: ncu_metrics.py <out.json> <key>=<report.ncu-rep>:<rays_per_launch>[:<note>] ..."""
import csv
import io
import json
import subprocess
import sys

out_path = sys.argv[1]
try:
    out = json.load(open(out_path))
except Exception:
    out = {}
for spec in sys.argv[2:]:
    key, rest = spec.split("=", 1)
    parts = rest.split(":", 2)
    rep, rays = parts[0], int(parts[1])
    note = parts[2] if len(parts) > 2 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[2]

    def val(name, scale_unit=True):
        i = hdr.index(name)
        v = float(r[i].replace(",", ""))
        u = units[i].lower()
        if scale_unit:  # ncu prints byte counts in the unit it likes best
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3,
                  "usecond": 1e-3, "msecond": 1, "nsecond": 1e-6, "second": 1e3}.get(u, 1)
        return v

    inst = val("smsp__inst_executed.sum", False)
    out[key] = {
        "kernel": r[hdr.index("Kernel Name")].split("(")[0].split("::")[-1],
        "rays_per_launch": rays,
        "warp_inst_per_ray": inst / rays,
        "lanes_per_inst": val("smsp__thread_inst_executed_per_inst_executed.ratio", False),
        "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
        "warps_eligible_per_cycle": val("smsp__warps_eligible.avg.per_cycle_active", False),
        "alu_pipe_pct": val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", False),
        "fma_pipe_pct": val("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", False),
        "l1_throughput_pct": val("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", False),
        "l1_hit_pct": val("l1tex__t_sector_hit_rate.pct", False),
        "l2_hit_pct": val("lts__t_sector_hit_rate.pct", False),
        "dram_bytes_read": val("dram__bytes_read.sum"),
        "dram_bytes_write": val("dram__bytes_write.sum"),
        "dram_throughput_pct": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False),
        "registers": val("launch__registers_per_thread", False),
        "ms_under_ncu": val("gpu__time_duration.sum"),
        "capture": (note + " " if note else "") + rep.split("/")[-1] + " (ncu --set full --clock-control none)",
    }
    print(key, json.dumps(out[key]))
json.dump(out, open(out_path, "w"), indent=1)
