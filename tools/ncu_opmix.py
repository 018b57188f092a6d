#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv` dump by SASS opcode and by stall reason, per kernel."""
import csv, sys, re, subprocess, collections
rep, pat = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kern, hdr = None, None
data = collections.OrderedDict()
for r in rows:
    if r and r[0] == "Kernel Name":
        kern = r[1]; data[kern] = []; hdr = None; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if kern and hdr and len(r) >= len(hdr) - 2:
        data[kern].append(dict(zip(hdr, r)))
for k, recs in data.items():
    if pat and pat not in k: continue
    print("==", k[:100])
    ops = collections.Counter(); stall = collections.Counter(); samples = 0
    tot = 0
    for d in recs:
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", d["Source"])
        op = m.group(1) if m else "?"
        base = op.split(".")[0]
        n = int(float(d["Instructions Executed"] or 0))
        ops[base] += n; tot += n
        for s in [h for h in d if h.startswith("stall_") and "Not Issued" not in h]:
            try: stall[s] += int(float(d[s] or 0))
            except ValueError: pass
    print(" total warp instr %.3e, static SASS lines %d" % (tot, len(recs)))
    for op, n in ops.most_common(28):
        print("   %-10s %6.2f%%  %.3e" % (op, 100.0 * n / tot, n))
    st = sum(stall.values())
    print(" stall samples:", ", ".join("%s %.1f%%" % (s.replace("stall_", ""), 100.0 * v / st) for s, v in stall.most_common(9)))
