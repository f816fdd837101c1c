#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of a kernel from an `ncu --set full --import-source on`
report:  python tools/srcprof.py <report.ncu-rep> <kernel> [top_n] [sort: inst|samp]"""
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    key = "# Samples" if len(sys.argv) > 4 and sys.argv[4] == "samp" else "Instructions Executed"
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
    src = {f: open(os.path.join(ROOT, "wgpu-cpu_b200", "csrc", f)).read().splitlines() for f in ("wgb_raster.cuh", "wgb_prelude.cuh")}
    tot = sum(num(d["Instructions Executed"]) for d in out)
    tots = sum(num(d["# Samples"]) for d in out)
    thr = sum(num(d["Thread Instructions Executed"]) for d in out)
    print(f"{kernel}: {tot / 1e9:.3f} G warp instructions, {thr / max(tot, 1):.1f} threads/instruction, {int(tots)} samples")
    for d in sorted(out, key=lambda d: -num(d[key]))[:topn]:
        f, n = d["file"], int(d["Line No"])
        text = src[f][n - 1].strip() if f in src and n <= len(src[f]) else d["Source"]
        print(f"{f[:11]:11s} L{n:>4d} inst={num(d['Instructions Executed']) / tot * 100:5.1f}% samp={num(d['# Samples']) / tots * 100:5.1f}% "
              f"thr={num(d['Thread Instructions Executed']) / max(num(d['Instructions Executed']), 1):4.1f} | {text[:95]}")


if __name__ == "__main__":
    main()
