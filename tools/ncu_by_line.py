#!/usr/bin/env python3
"""Aggregate an ncu SASS source-page CSV by CUDA source line, using nvdisasm -g line info.
usage: ncu_by_line.py <src.csv> <nvdisasm_lines.txt> <mangled-kernel-substring> [topN]"""
import csv, re, sys, collections
src_csv, dis, key = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# instruction index -> (file, line, inlined chain)
lines = []
cur = None
infn = False
for ln in open(dis):
    if ln.startswith(".text."):
        infn = key in ln
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", ln):
        lines.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr, data = rows[1], rows[2:]
ci = hdr.index("Instructions Executed"); ct = hdr.index("Thread Instructions Executed")
cs = hdr.index("# Samples")
assert len(data) == len(lines), (len(data), len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for r, l in zip(data, lines):
    a = agg[l]
    a[0] += int(r[ci]); a[1] += int(r[ct]); a[2] += int(r[cs]); a[3] += 1
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
print("total warp-inst %.3g, samples %d" % (tot, tots))
print("%-28s %6s %9s %6s %6s %7s" % ("file:line", "ninst", "warpinst", "share", "thr", "smp%"))
for l, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:topn]:
    print("%-28s %6d %8.1fM %5.1f%% %6.1f %6.1f%%" % ("%s:%d" % l, a[3], a[0] / 1e6, 100.0 * a[0] / tot, a[1] / max(a[0], 1), 100.0 * a[2] / tots))
