#!/usr/bin/env python3
"""Aggregate an ncu SASS source-page CSV by named source regions (function line ranges).
usage: ncu_by_region.py <src.csv> <nvdisasm_lines.txt> <mangled-kernel-substring>"""
import csv, re, sys, collections
src_csv, dis, key = sys.argv[1], sys.argv[2], sys.argv[3]
REG = [
    ("vode_core.cuh", 61, 83, "pick helpers"),
    ("vode_core.cuh", 84, 140, "dvset_bdf"),
    ("vode_core.cuh", 141, 176, "dvjust_coeffs"),
    ("vode_core.cuh", 215, 239, "dewset/norm_y0"),
    ("vode_core.cuh", 240, 249, "predict"),
    ("vode_core.cuh", 250, 259, "retract"),
    ("vode_core.cuh", 260, 274, "rescale"),
    ("vode_core.cuh", 275, 287, "get_col"),
    ("vode_core.cuh", 288, 317, "order_change"),
    ("vode_core.cuh", 318, 334, "dvindy0"),
    ("vode_core.cuh", 375, 483, "dvhin/init"),
    ("vode_core.cuh", 484, 594, "dvnlsd"),
    ("vode_core.cuh", 595, 760, "run: head..dsm"),
    ("vode_core.cuh", 760, 814, "run: failure paths"),
    ("vode_core.cuh", 815, 836, "run: accept update"),
    ("vode_core.cuh", 836, 943, "run: order/eta select"),
    ("vode_core.cuh", 943, 975, "run: tail"),
    ("bvode_math.cuh", 25, 45, "powi"),
    ("bvode_math.cuh", 110, 125, "rpow/recip_int"),
    ("bvode_math.cuh", 126, 162, "log2_fast"),
    ("bvode_math.cuh", 163, 190, "exp2_fast"),
    ("policy_thread.cuh", 30, 45, "GlobalVec/Tol"),
    ("policy_thread.cuh", 120, 180, "norm/msq/etc"),
    ("policy_thread.cuh", 181, 233, "dgefa"),
    ("policy_thread.cuh", 234, 266, "dgesl"),
    ("policy_thread.cuh", 267, 348, "jac_setup"),
    ("problems.cuh", 0, 999, "f/jac"),
    ("bvode_kernels.cuh", 0, 999, "kernel shell"),
]
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
