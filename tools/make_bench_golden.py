#!/usr/bin/env python
"""Writes tests/golden/bench_frames_sha256.json: SHA-256 of the colour bytes (and depth bits) the CPU oracle renders for
the BASELINE.json configurations AT THEIR FULL SIZE -- the frames `bench.py` renders.  bench.py hashes the frame it has
assembled on rank 0 after the timed region (at every N, for both presenters) and compares it with these digests, so the
driver-run BENCH / SCALE records carry hardware parity for the multi-GPU path ("results are bit-identical
to 1 GPU"; serial semantics state.rs:519-593).

    python tools/make_bench_golden.py            # ~2 minutes of CPU; regenerate after a deliberate change of the oracle
"""
import hashlib
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

C5_FRAMES = 64


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def generate(only=None) -> dict:
    from oracle import pyoracle
    from wgpu_cpu_b200 import scenes as S
    import bench
    out = {}
    for key, (workload, make) in bench.CONFIGS.items():
        if only and key not in only:
            continue
        if key == "c5":
            frames = []
            for f in range(C5_FRAMES):
                fr = pyoracle.render(S.hello_texture(3840, 2160, yaw=f * 2.0 * math.pi / C5_FRAMES), want_coverage=False)
                frames.append(sha(fr.color))
            out[key] = {"workload": workload, "frames": frames, "batch": hashlib.sha256("".join(frames).encode()).hexdigest()}
            continue
        fr = pyoracle.render(make(S), want_coverage=False)
        out[key] = {"workload": workload, "color": sha(fr.color)}
        if fr.depth is not None:
            out[key]["depth"] = sha(fr.depth.view(np.uint32))
    return out


if __name__ == "__main__":
    path = os.path.join(ROOT, "tests", "golden", "bench_frames_sha256.json")
    old = json.load(open(path)) if os.path.exists(path) else {}
    old.update(generate(sys.argv[1:] or None))
    with open(path, "w") as f:
        json.dump(old, f, indent=1, sort_keys=True)
        f.write("\n")
    print(path)
