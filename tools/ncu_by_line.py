#!/usr/bin/env python3
"""Dynamic SASS census per CUDA source line: joins `ncu --page source --csv` (executed counts per SASS instruction)
with `nvdisasm --print-line-info` of the same cubin (source line per SASS instruction) by instruction order.
usage: ncu_by_line.py report.ncu-rep dump.dis <ncu kernel substring> <mangled substring> [top]"""
import csv, re, sys, subprocess, collections
rep, dis, kpat, mpat = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kern = None; hdr = None; recs = []
for r in rows:
    if r and r[0] == "Kernel Name": kern = r[1]; hdr = None; continue
    if r and r[0] == "Address": hdr = r; continue
    if kern and kpat in kern and hdr and len(r) >= len(hdr) - 2: recs.append(dict(zip(hdr, r)))
lines = []; on = False; cur = None
for line in open(dis):
    if line.startswith(".text."): on = mpat in line; continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m: lines.append((cur, m.group(1).split(".")[0]))
print("ncu instr", len(recs), "nvdisasm instr", len(lines))
n = min(len(recs), len(lines))
cnt = collections.Counter(); f64 = collections.Counter(); ops = collections.defaultdict(collections.Counter); lsb = collections.Counter()
tot = 0
for i in range(n):
    e = int(float(recs[i]["Instructions Executed"] or 0))
    cur, op = lines[i]
    cnt[cur] += e; tot += e; ops[cur][op] += e
    if op in ("DADD", "DMUL", "DFMA", "DSETP"): f64[cur] += e
    lsb[cur] += int(float(recs[i].get("# Samples") or 0))
allsm = sum(lsb.values())
print("total warp instr %.3e  fp64 %.3e" % (tot, sum(f64.values())))
src = {}
for (f, l), e in cnt.most_common(top):
    if f not in src:
        try: src[f] = open("/root/repo/cholla_b200/csrc/" + f).read().splitlines()
        except Exception: src[f] = []
    text = src[f][l - 1].strip()[:80] if l - 1 < len(src[f]) else ""
    o = " ".join("%s %.0f%%" % (k, 100.0 * v / e) for k, v in ops[(f, l)].most_common(4))
    print("%5.1f%% (f64 %4.1f%%, samples %4.1f%%) %s:%d  %s\n        [%s]" % (100.0 * e / tot, 100.0 * f64[(f, l)] / tot, 100.0 * lsb[(f, l)] / max(allsm, 1), f, l, text, o))
