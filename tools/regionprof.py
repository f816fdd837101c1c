#!/usr/bin/env python
"""Per-function instruction shares of a kernel from an `ncu --set full --import-source on` report
(inlined code is attributed to the function whose source lines it came from):
python tools/regionprof.py <report.ncu-rep> <kernel>"""
import csv
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", kernel],
                         capture_output=True, text=True).stdout
    fname, hdr, out = None, None, []
    for r in csv.reader(txt.splitlines()):
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0].isdigit():
            d = {}
            for k, v in zip(hdr, r):
                d.setdefault(k, v)
            d["file"] = fname
            out.append(d)
    num = lambda x: float(x) if x.replace(".", "").isdigit() else 0.0
    starts = {}
    for f in ("wgb_raster.cuh", "wgb_prelude.cuh"):
        lines = open(os.path.join(ROOT, "wgpu-cpu_b200", "csrc", f)).read().splitlines()
        cur = []
        for i, l in enumerate(lines, 1):
            m = re.match(r"^(?:template.*>\s*)?(?:WGB_DEV|__global__|static|__device__|extern).*?\b(\w+)\s*\(", l)
            if m and not l.startswith(" "):
                cur.append((i, m.group(1)))
        starts[f] = cur
    def region(f, n):
        name = "?"
        for i, nm in starts.get(f, []):
            if i <= n:
                name = nm
            else:
                break
        return f"{f.split('.')[0][4:]}:{name}"
    agg = defaultdict(lambda: [0.0, 0.0, 0.0])
    for d in out:
        a = agg[region(d["file"], int(d["Line No"]))]
        a[0] += num(d["Instructions Executed"]); a[1] += num(d["Thread Instructions Executed"]); a[2] += num(d["# Samples"])
    tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
    print(f"{kernel}: {tot / 1e9:.3f} G warp instructions")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if a[0] / tot < 0.002:
            continue
        print(f"{k:45s} inst={a[0] / tot * 100:5.1f}% samp={a[2] / max(tots, 1) * 100:5.1f}% thr/inst={a[1] / max(a[0], 1):4.1f}")


if __name__ == "__main__":
    main()
