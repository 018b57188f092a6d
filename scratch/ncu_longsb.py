#!/usr/bin/env python3
"""Which SASS instructions wait on the long scoreboard: for each kernel in an `ncu --page source --csv` dump, find the
producer (the most recent LDG/LDL/LDS/... writing a register the stalled instruction reads is not tracked; instead
list the stalled instruction and the nearest preceding memory op) -- good enough to tell spill reloads from global loads.
usage: ncu_longsb.py dump.csv kernel-substring"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
pat = sys.argv[2]
kern = None; hdr = None; data = collections.OrderedDict()
for r in rows:
    if r and r[0] == "Kernel Name": kern = r[1]; data[kern] = []; hdr = None; continue
    if r and r[0] == "Address": hdr = r; continue
    if kern and hdr and len(r) >= len(hdr) - 2: data[kern].append(dict(zip(hdr, r)))
for k, recs in data.items():
    if pat not in k: continue
    print("==", k[:110])
    tot = sum(int(float(d["stall_long_sb"] or 0)) for d in recs)
    allsm = sum(int(float(d["# Samples"] or 0)) for d in recs)
    print("long_sb samples", tot, "of", allsm)
    # attribute each stalled instruction to the kind of memory op whose result it most plausibly waits for:
    # scan back for the nearest memory load that writes one of its source registers
    def regs(s):
        return set(re.findall(r"\bR(\d+)\b", s))
    by = collections.Counter()
    memops = collections.Counter()
    for i, d in enumerate(recs):
        n = int(float(d["stall_long_sb"] or 0))
        src = d["Source"]
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
        op = m.group(1).split(".")[0] if m else "?"
        if op in ("LDG", "LDL", "LDS", "STL", "STG", "STS", "LDGSTS", "ATOMG", "LDC"):
            memops[op] += int(float(d["Instructions Executed"] or 0))
        if n == 0: continue
        need = regs(src.split(",", 1)[1]) if "," in src else regs(src)
        found = "?"
        for j in range(i - 1, max(i - 400, -1), -1):
            s2 = recs[j]["Source"]
            m2 = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s+R(\d+)", s2)
            if not m2: continue
            op2 = m2.group(1).split(".")[0]
            if op2 in ("LDG", "LDL", "LDS", "LDC", "MUFU", "SHFL", "ATOMG", "S2R"):
                dst = int(m2.group(2))
                w = 2 if (".64" in m2.group(1)) else 1
                if any(str(dst + q) in need for q in range(w)):
                    found = op2; break
        by[found] += n
    for a, b in by.most_common(): print("  waits on %-6s %6d  %.1f%%" % (a, b, 100.0 * b / max(tot, 1)))
    print("  dynamic memory ops:", dict(memops))
