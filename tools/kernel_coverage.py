#!/usr/bin/env python
"""Line coverage of the CUDA kernel source (wgb_raster.cuh, wgb_prelude.cuh) by the GPU test-suite, measured on the
software model of tests/cusim: every pipeline translation unit is compiled by g++ with --coverage, the suite runs, and
the per-pipeline gcov reports are merged (a line counts as executed if any pipeline variant executed it).

    python tools/kernel_coverage.py [--keep DIR]        # about 3 minutes on 8 cores

Prints the lines of the kernel source that some variant compiled and no test executed."""
import collections
import os
import re
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARDWARE_ONLY = ["tests/test_parity_gpu.py::test_pinned_uploads_on_the_copy_stream_are_ordered_with_rendering"]


def main():
    keep = sys.argv[sys.argv.index("--keep") + 1] if "--keep" in sys.argv else None
    cache = keep or tempfile.mkdtemp(prefix="wgb_kernel_cov_")
    env = dict(os.environ, WGB_CUSIM="1", CUSIM_KEEP="1", CUSIM_OPT="-O0 --coverage", CUSIM_CACHE=cache)
    cmd = [sys.executable, "-m", "pytest", "tests", "-q", "-m", "gpu", "-p", "no:cacheprovider", "-n", str(min(8, os.cpu_count() or 1)),
           "-k", "not full_size and not peer_presenter"]
    for t in HARDWARE_ONLY:
        cmd += ["--deselect", t]
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True)
    print(p.stdout.strip().splitlines()[-1])
    files = ("wgb_raster.cuh", "wgb_prelude.cuh")
    agg = {f: collections.defaultdict(lambda: [0, False]) for f in files}
    text = {}
    for d in sorted(os.listdir(cache)):
        d = os.path.join(cache, d)
        if not os.path.isfile(os.path.join(d, "out.so-tu.gcda")):
            continue
        for ext in ("gcda", "gcno"):
            shutil.copy(os.path.join(d, f"out.so-tu.{ext}"), os.path.join(d, f"tu.{ext}"))
        subprocess.run(["gcov", "tu.cpp"], cwd=d, capture_output=True)
        for f in files:
            report = os.path.join(d, f + ".gcov")
            if not os.path.exists(report):
                continue
            for line in open(report, errors="replace"):
                m = re.match(r"\s*([^:]+):\s*(\d+):(.*)", line)
                if not m or m.group(2) == "0" or m.group(1).strip() == "-":
                    continue
                count, number = m.group(1).strip(), int(m.group(2))
                text[(f, number)] = m.group(3).rstrip()
                entry = agg[f][number]
                entry[1] = True
                if count not in ("#####", "=====", "$$$$$"):
                    entry[0] += int(count.rstrip("*"))
    for f in files:
        compiled = [n for n, e in agg[f].items() if e[1]]
        missed = sorted(n for n in compiled if agg[f][n][0] == 0)
        print(f"{f}: {len(compiled) - len(missed)} of {len(compiled)} compiled lines executed ({100.0 * (len(compiled) - len(missed)) / max(len(compiled), 1):.1f} %)")
        for n in missed:
            print(f"    {n:5d}: {text[(f, n)][:140]}")
    if not keep:
        shutil.rmtree(cache, ignore_errors=True)


if __name__ == "__main__":
    main()
