#!/usr/bin/env python3
"""Extract the reference's golden vectors into tests/golden/*.json.

Reads (read-only) /root/reference/test/dvdemo.txt -- the expected transcript of
the reference's dvdemo program (24 runs) -- and the trailing comment of
/root/reference/test/example.f90 (Robertson, Cray-1 output).  Only numbers are
extracted; they are kept as the printed strings so that tests can compare the
oracle's output digit for digit after formatting with the same edit descriptors
(dvdemo.f90: 99003 FORMAT(1X,D15.5,D16.5,D14.3,I5,D14.3), 99006
FORMAT(1X,D15.5,D14.3,I5,D14.3), 99009 statistics).

Run in the build container (the GPU box has no /root/reference):
    python tools/make_golden.py
"""
import json
import os
import re

REF = "/root/reference/test"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def parse_dvdemo(path):
    runs = []
    problem = 0
    cur = None
    with open(path) as fh:
        lines = fh.read().splitlines()
    for ln in lines:
        if ln.startswith(" Problem 1"):
            problem = 1
        elif ln.startswith(" Problem 2"):
            problem = 2
        m = re.match(r"\s*Solution with MF =\s*(-?\d+)", ln)
        if m:
            cur = {"problem": problem, "mf": int(m.group(1)), "rows": [], "stats": {}}
            runs.append(cur)
            continue
        if cur is None:
            continue
        toks = ln.split()
        if toks and re.match(r"^-?0\.\d+D[+-]\d\d$", toks[0]) and len(toks) in (4, 5):
            cur["rows"].append(toks)
            continue
        m = re.match(r"\s*RWORK size =\s*(\d+)\s+IWORK size =\s*(\d+)", ln)
        if m:
            cur["stats"]["lenrw"] = int(m.group(1))
            cur["stats"]["leniw"] = int(m.group(2))
        for key, pat in (("nst", r"Number of steps ="), ("nfe", r"Number of f-s\s+="),
                         ("nfea", r"\(excluding J-s\) ="), ("nje", r"Number of J-s\s+="),
                         ("nlu", r"Number of LU-s\s+="),
                         ("nni", r"Number of nonlinear iterations ="),
                         ("ncfn", r"Number of nonlinear convergence failures ="),
                         ("netf", r"Number of error test failures =")):
            m = re.match(r"\s*" + pat + r"\s*(\d+)", ln)
            if m:
                cur["stats"][key] = int(m.group(1))
        m = re.match(r"\s*Error overrun =\s*(\S+)", ln)
        if m:
            cur["overrun"] = m.group(1)
    m = re.search(r"Number of errors encountered =\s*(\d+)", "\n".join(lines))
    return {"source": "reference test/dvdemo.txt", "nerr": int(m.group(1)), "runs": runs}


def parse_example(path):
    rows, stats = [], {}
    with open(path) as fh:
        for ln in fh:
            m = re.match(r"!\s*at t =\s*(\S+)\s+y =\s*(\S+)\s+(\S+)\s+(\S+)", ln)
            if m:
                rows.append([float(x) for x in m.groups()])
            for key, pat in (("nst", r"no\. steps ="), ("nfe", r"no\. f-s ="), ("nje", r"no\. j-s ="),
                             ("nlu", r"no\. lu-s ="), ("nni", r"no\. nonlinear iterations ="),
                             ("ncfn", r"no\. nonlinear convergence failures ="),
                             ("netf", r"no\. error test failures =")):
                m = re.search(pat + r"\s*(\d+)", ln)
                if m and ln.lstrip().startswith("!"):
                    stats[key] = int(m.group(1))
    return {"source": "reference test/example.f90 trailing comment (Cray-1, CFT: non-IEEE)",
            "rows": rows, "stats": stats}


def main():
    os.makedirs(OUT, exist_ok=True)
    d = parse_dvdemo(os.path.join(REF, "dvdemo.txt"))
    assert len(d["runs"]) == 24, len(d["runs"])
    with open(os.path.join(OUT, "dvdemo_golden.json"), "w") as fh:
        json.dump(d, fh, indent=1)
    e = parse_example(os.path.join(REF, "example.f90"))
    assert len(e["rows"]) == 12 and len(e["stats"]) == 7, e
    with open(os.path.join(OUT, "robertson_example_golden.json"), "w") as fh:
        json.dump(e, fh, indent=1)
    print("wrote", len(d["runs"]), "dvdemo runs and the Robertson example")


if __name__ == "__main__":
    main()
