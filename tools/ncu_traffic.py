#!/usr/bin/env python3
"""profiles/roofline_traffic.json from `ncu --set full --import-source on` captures of one step's sweeps (one report per
arithmetic mode): per launch and per step DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum), durations, fp64-pipe
and issue-slot utilisation, spill requests, and -- from the source page -- the dynamic warp-instruction counts (all / fp64
pipe) that give the instructions per cell-update behind bench.py's `roofline.secondary`.
usage: ncu_traffic.py <cells> <commit> mode=report.ncu-rep [mode=report ...] > profiles/roofline_traffic.json"""
import collections
import csv
import json
import re
import subprocess
import sys

UN = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TU = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}
FP64 = {"DADD", "DMUL", "DFMA", "DSETP", "DMNMX"}
cells, commit = int(sys.argv[1]), sys.argv[2]
out = {"_how": "ncu --set full --clock-control none --import-source on, one step's direction sweeps of `python bench.py` "
               "(512^3 blast PPMC+HLLC fp64) after 3 warm-up steps; per-launch values are cold-cache and serialised; "
               "instruction counts are warp instructions x 32 / cells", "commit": commit}
for arg in sys.argv[3:]:
    mode, rep = arg.split("=", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, k, table=None):
        v = float(r[col[k]].replace(",", ""))
        return v * table[units[col[k]]] if table else v

    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    per_kernel, kern, h = collections.OrderedDict(), None, None
    for r in csv.reader(src.splitlines()):
        if r and r[0] == "Kernel Name":
            kern = len(per_kernel)
            per_kernel[kern] = [0, 0]
            h = None
            continue
        if r and r[0] == "Address":
            h = r
            continue
        if kern is not None and h and len(r) >= len(h) - 2:
            d = dict(zip(h, r))
            m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", d.get("Source", ""))
            n = int(float(d.get("Instructions Executed") or 0))
            per_kernel[kern][0] += n
            if m and m.group(1) in FP64:
                per_kernel[kern][1] += n
    launches, tot = [], collections.Counter()
    for i, r in enumerate(data):
        b = val(r, "dram__bytes_read.sum", UN) + val(r, "dram__bytes_write.sum", UN)
        ms = val(r, "gpu__time_duration.sum", TU)
        f64 = val(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
        iss = val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
        loc = val(r, "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum") + val(r, "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum")
        # the source page lists every launch twice (two views of the same SASS): launch i is section 2 i
        sec = per_kernel.get(2 * i if len(per_kernel) == 2 * len(data) else i)
        winst, wfp64 = sec if sec else [val(r, "smsp__inst_executed.sum"), 0]
        launches.append({"kernel": r[col["Kernel Name"]][:60], "dram_bytes": b, "ms": ms, "fp64_pipe_busy_pct": f64,
                         "issue_busy_pct": iss, "dram_gbs": b / ms / 1e6, "warp_inst": winst, "warp_inst_fp64": wfp64,
                         "local_requests": loc, "regs": val(r, "launch__registers_per_thread")})
        tot["b"] += b
        tot["ms"] += ms
        tot["f64"] += f64 * ms
        tot["iss"] += iss * ms
        tot["winst"] += winst
        tot["wfp64"] += wfp64
        tot["loc"] += loc
    out[mode] = {"dram_bytes_per_step": tot["b"], "ms_sum_under_ncu": tot["ms"], "fp64_pipe_busy_pct": tot["f64"] / tot["ms"],
                 "issue_busy_pct": tot["iss"] / tot["ms"], "cells": cells, "dram_bytes_per_cell_update": tot["b"] / cells,
                 "inst_per_cell": tot["winst"] * 32.0 / cells, "fp64_inst_per_cell": tot["wfp64"] * 32.0 / cells,
                 "local_requests_per_step": tot["loc"], "commit": commit, "launches": launches}
print(json.dumps(out, indent=1))
