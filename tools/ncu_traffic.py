#!/usr/bin/env python3
"""profiles/ncu_traffic.json: measured DRAM bytes per system of the three hot kernels, from the committed summaries
of this round's `ncu --set full` captures (bench.py reads it for `roofline.traffic`).
usage: python tools/ncu_traffic.py <tag>      e.g. r2o -> profiles/r02_ncu_thread_kernel_r2o.txt, ..._warp_advection_r2o.txt, ..._warp_mech32_r2o.txt"""
import json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
prefix = os.environ.get("PREFIX", "r02")
files = {"robertson": "%s_ncu_thread_kernel_%s.txt" % (prefix, tag),
         "advection": "%s_ncu_warp_advection_%s.txt" % (prefix, tag),
         "mech32": "%s_ncu_warp_mech32_%s.txt" % (prefix, tag)}
out = {}
for key, fn in files.items():
    path = os.path.join(ROOT, "profiles", fn)
    if not os.path.exists(path):
        continue
    txt = open(path).read()
    def val(name):
        m = re.search(r"^\s*%s\s+([0-9.]+)\s+(\S+)" % re.escape(name), txt, re.M)
        v, unit = float(m.group(1)), m.group(2).lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
    grid = int(float(re.search(r"launch__grid_size\s+([0-9.]+)", txt).group(1)))
    block = int(float(re.search(r"launch__block_size\s+([0-9.]+)", txt).group(1)))
    nsys = grid * block if key == "robertson" else grid * (block // 32)  # one system per thread / per warp
    b = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    out[key] = {"bytes_per_system": round(b / nsys, 1), "source": "profiles/%s (%.1f MB for %d systems)" % (fn, b / 1e6, nsys)}
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
