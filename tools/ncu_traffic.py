#!/usr/bin/env python3
"""profiles/roofline_traffic.json from `ncu --set full` captures of one step's six sweeps (one report per arithmetic mode):
DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch and per step, durations, fp64-pipe utilisation.
usage: ncu_traffic.py mode=report.ncu-rep [mode=report ...] > profiles/roofline_traffic.json"""
import csv, json, subprocess, sys
UN = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TU = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}
out = {"_how": "ncu --set full --clock-control none, one step (six vl_dir_kernel launches) of bench.py at 512^3 PPMC+HLLC fp64 "
               "after 3 warm-up steps; per-launch values are cold-cache and serialised"}
for arg in sys.argv[1:]:
    mode, rep = arg.split("=", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    def val(r, k, table):
        return float(r[col[k]].replace(",", "")) * table[units[col[k]]]
    launches, tot_b, tot_ms, w = [], 0.0, 0.0, 0.0
    for r in data:
        b = val(r, "dram__bytes_read.sum", UN) + val(r, "dram__bytes_write.sum", UN)
        ms = val(r, "gpu__time_duration.sum", TU)
        f64 = float(r[col["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]])
        launches.append({"kernel": r[col["Kernel Name"]][:60], "dram_bytes": b, "ms": ms, "fp64_pipe_busy_pct": f64,
                         "dram_gbs": b / ms / 1e6})
        tot_b += b; tot_ms += ms; w += f64 * ms
    out[mode] = {"dram_bytes_per_step": tot_b, "ms_sum_under_ncu": tot_ms, "fp64_pipe_busy_pct": w / tot_ms,
                 "cells": 512 ** 3, "dram_bytes_per_cell_update": tot_b / 512 ** 3, "launches": launches}
print(json.dumps(out, indent=1))
