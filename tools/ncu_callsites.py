#!/usr/bin/env python3
"""Dynamic counts of one SASS opcode by inlined call chain.
usage: ncu_callsites.py <src.csv> <nvdisasm -gi output> <mangled-kernel-substring> <opcode substring> [depth]"""
import csv, re, sys, collections
src_csv, dis, key, opc = sys.argv[1:5]
depth = int(sys.argv[5]) if len(sys.argv) > 5 else 3
entries = []; cur = None; infn = False; prev_dir = False
for ln in open(dis):
    if ln.startswith(".text."):
        infn = key in ln; prev_dir = False; continue
    if not infn: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        if not prev_dir: cur = [(m.group(1).split("/")[-1], int(m.group(2)))]  # innermost frame first
        m2 = re.search(r'inlined at "([^"]+)", line (\d+)', ln)
        if m2: cur.append((m2.group(1).split("/")[-1], int(m2.group(2))))
        prev_dir = True
        continue
    prev_dir = False
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", ln): entries.append(cur)
rows = list(csv.reader(open(src_csv))); hdr, data = rows[1], rows[2:]
ci = hdr.index("Instructions Executed"); so = hdr.index("Source"); cs = hdr.index("# Samples"); ct = hdr.index("Thread Instructions Executed")
assert len(entries) == len(data), (len(entries), len(data))
agg = collections.defaultdict(lambda: [0, 0, 0])
for r, l in zip(data, entries):
    if opc in r[so]:
        a = agg[tuple(l[:depth])]; a[0] += int(r[ci]); a[1] += int(r[cs]); a[2] += int(r[ct])
tot = sum(a[0] for a in agg.values())
print("total %.1fM" % (tot / 1e6))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print("%7.1fM thr %4.1f smp %6d  %s" % (a[0] / 1e6, a[2] / max(a[0], 1), a[1], " <- ".join("%s:%d" % x for x in k)))
