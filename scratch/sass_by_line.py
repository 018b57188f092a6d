#!/usr/bin/env python3
"""Static SASS census per CUDA source line for one kernel: nvdisasm --print-line-info output -> instructions per
(file:line), split into fp64-pipe / other.  usage: sass_by_line.py dump.dis <mangled-substring> [top]"""
import re, sys, collections
dis, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
on = False; cur = None
cnt = collections.Counter(); f64 = collections.Counter(); ops = collections.defaultdict(collections.Counter)
tot = 0
for line in open(dis):
    if line.startswith(".text."):
        on = pat in line; continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1).split(".")[0]
        cnt[cur] += 1; tot += 1
        ops[cur][op] += 1
        if op in ("DADD", "DMUL", "DFMA", "DSETP"): f64[cur] += 1
print("total SASS", tot, "fp64", sum(f64.values()))
src = {}
for (f, l), n in cnt.most_common(top):
    if f not in src:
        try: src[f] = open("/root/repo/cholla_b200/csrc/" + f).read().splitlines()
        except Exception: src[f] = []
    text = src[f][l - 1].strip()[:90] if l - 1 < len(src[f]) else ""
    o = " ".join("%s%d" % (k, v) for k, v in ops[(f, l)].most_common(5))
    print("%4d (f64 %3d) %s:%d  %s\n        [%s]" % (n, f64[(f, l)], f, l, text, o))
