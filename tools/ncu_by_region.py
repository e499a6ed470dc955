#!/usr/bin/env python3
"""Aggregate an ncu SASS source-page CSV by named source regions (function line ranges).
usage: ncu_by_region.py <src.csv> <nvdisasm_lines.txt> <mangled-kernel-substring>"""
import csv, re, sys, collections
src_csv, dis, key = sys.argv[1], sys.argv[2], sys.argv[3]
import os
def _regions():
    """(file, first line, last line, name) of every device function in dvode_b200/csrc, plus the
    marked sections of Engine::run ('// ----' comments)."""
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dvode_b200", "csrc")
    out = []
    for fn in sorted(os.listdir(root)):
        if not fn.endswith((".cuh", ".cu", ".h")): continue
        starts = []
        inrun = False
        for n, ln in enumerate(open(os.path.join(root, fn)), 1):
            m = re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:BV_DEV|__device__|__global__)[^;(]*?\b([A-Za-z_]\w*)\s*\(", ln)
            if m and not ln.strip().startswith("//"):
                name = m.group(1)
                if name in ("__launch_bounds__", "__noinline__"):
                    m2 = re.search(r"\b([A-Za-z_]\w*)\s*\($", ln.rstrip().rstrip("{").rstrip())
                    continue
                starts.append((n, name)); inrun = (name == "run")
                continue
            if inrun:
                m = re.match(r"\s*// ---- (.{3,40})", ln)
                if m: starts.append((n, "run/" + m.group(1).strip()[:34]))
        for k, (n, name) in enumerate(starts):
            end = starts[k + 1][0] - 1 if k + 1 < len(starts) else 10 ** 9
            out.append((fn, n, end, name))
    return out
REG = _regions()
lines = []; cur = None; infn = False
for ln in open(dis):
    if ln.startswith(".text."):
        infn = key in ln; continue
    if not infn: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", ln): lines.append(cur)
rows = list(csv.reader(open(src_csv))); hdr, data = rows[1], rows[2:]
ci = hdr.index("Instructions Executed"); ct = hdr.index("Thread Instructions Executed"); cs = hdr.index("# Samples")
so = hdr.index("Source")
assert len(data) == len(lines), (len(data), len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
def region(l):
    if l is None: return "?"
    for f, a, b, n in REG:
        if l[0] == f and a <= l[1] <= b: return n
    return "%s:%d" % l
for r, l in zip(data, lines):
    a = agg[region(l)]
    a[0] += int(r[ci]); a[1] += int(r[ct]); a[2] += int(r[cs]); a[3] += 1
    op = r[so].split()[0] if not r[so].strip().startswith("@") else r[so].split()[1]
    if op.startswith(("DFMA", "DMUL", "DADD", "DSETP", "MUFU")): a[4] += int(r[ci])
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
print("total warp-inst %.3g, samples %d, fp64-pipe warp-inst %.3g" % (tot, tots, sum(a[4] for a in agg.values())))
print("%-26s %6s %9s %6s %6s %7s %7s" % ("region", "ninst", "warpinst", "share", "thr", "smp%", "fp64%"))
for l, a in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    print("%-26s %6d %8.1fM %5.1f%% %6.1f %6.1f%% %6.1f%%" % (l, a[3], a[0] / 1e6, 100.0 * a[0] / tot, a[1] / max(a[0], 1), 100.0 * a[2] / tots, 100.0 * a[4] / max(a[0], 1)))
