#!/usr/bin/env python3
"""Static instruction count of one kernel by source region (no GPU needed).
usage: sass_by_region.py <object.o> <mangled-kernel-substring>"""
import re, subprocess, sys, collections, tempfile, os, glob
obj, key = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cubin = glob.glob(os.path.join(d, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
REG = None
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_by_region.py")).read()
exec(src[src.index("import os\ndef _regions"):src.index("lines = []")])
agg = collections.Counter(); byline = collections.Counter(); cur = None; infn = False; n = 0
for ln in dis.splitlines():
    if ln.startswith(".text."):
        infn = key in ln; continue
    if not infn: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", ln):
        n += 1
        r = "?"
        if cur:
            r = "%s:%d" % cur
            for f, a, b, name in REG:
                if cur[0] == f and a <= cur[1] <= b: r = name; break
        agg[r] += 1; byline[cur] += 1
print("total", n)
for r, c in agg.most_common(): print("%-26s %5d" % (r, c))
if len(sys.argv) > 3:
    for l, c in byline.most_common(int(sys.argv[3])): print(l, c)
