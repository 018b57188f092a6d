#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md quotes: per-kernel duration, DRAM
bytes, pipe utilisation, occupancy, stall reasons and the dynamic SASS opcode mix.  Usage:
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
KEYS = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs/thread"), ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of 64/SM"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe busy %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "L1 hit % (global ld)"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "local (spill) ld requests"),
        ("l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "local (spill) st requests")]
STALLS = ["wait", "long_scoreboard", "barrier", "math_pipe_throttle", "no_instructions", "short_scoreboard", "selected",
          "not_selected", "branch_resolving", "dispatch_stall", "lg_throttle", "mio_throttle"]
print("# ncu --set full summary: `%s`\n" % rep.split("/")[-1])
print("Captured with `ncu --set full --clock-control none --import-source on` under gpurun on one B200; per-launch\n"
      "values are cold-cache and serialised (shares, not absolutes, are comparable with the CUDA-event timings).\n")
for r in data:
    name = r[col["Kernel Name"]]
    print("## `%s`\n" % name[:160])
    print("| metric | value |\n|---|---|")
    for k, label in KEYS:
        if k in col:
            print("| %s | %s %s |" % (label, r[col[k]], units[col[k]]))
    tot = 0
    st = {}
    for s in STALLS:
        k = "smsp__pcsamp_warps_issue_stalled_" + s
        if k in col:
            st[s] = float(r[col[k]].replace(",", ""))
            tot += st[s]
    if tot:
        print("| stall samples (share) | %s |" % ", ".join("%s %.0f%%" % (s, 100 * v / tot) for s, v in
                                                           sorted(st.items(), key=lambda kv: -kv[1]) if v / tot > 0.01))
    print()
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern, h = None, None
recs = collections.OrderedDict()
for r in csv.reader(src.splitlines()):
    if r and r[0] == "Kernel Name":
        kern = r[1]
        recs[kern] = []
        h = None
    elif r and r[0] == "Address":
        h = r
    elif kern and h and len(r) >= len(h) - 2:
        recs[kern].append(dict(zip(h, r)))
print("## dynamic SASS opcode mix (warp instructions executed, from the source page)\n")
for k, rs in recs.items():
    ops, tot = collections.Counter(), 0
    for d in rs:
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", d["Source"])
        op = (m.group(1) if m else "?").split(".")[0]
        n = int(float(d["Instructions Executed"] or 0))
        ops[op] += n
        tot += n
    if not tot:
        continue
    f64 = sum(v for o, v in ops.items() if o in ("DFMA", "DADD", "DMUL", "DSETP"))
    print("`%s`: %.3e warp instructions, %d static SASS lines, fp64-pipe share %.1f%%  " % (k[:120], tot, len(rs), 100.0 * f64 / tot))
    print(", ".join("%s %.1f%%" % (o, 100.0 * n / tot) for o, n in ops.most_common(14)) + "\n")
