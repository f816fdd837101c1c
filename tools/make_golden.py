#!/usr/bin/env python
"""Writes tests/golden/oracle_frames_sha256.json: SHA-256 of the colour bytes and depth bits the CPU oracle renders for a fixed
set of scenes (every BASELINE shape at test size, the parity shaders, a few fuzz seeds).

The reference's own golden images are git-LFS pointers, and the reference cannot be built here, so these
are not reference outputs: they pin the ORACLE -- the frames recorded here are the ones the B200 path was bit-exact
against in round 1 (and the software model of tests/cusim since), so an accidental change of the oracle's arithmetic
shows up as a changed hash instead of silently moving the target the CUDA path is compared with.

    python tools/make_golden.py            # regenerate after a deliberate change of the oracle"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def scenes():
    from wgpu_cpu_b200 import scenes as S
    out = [S.colored_triangle(v, 128, 128) for v in ("default", "cull_front", "draw_backwards", "draw_backwards_no_cull", "lines")]
    out += [S.hello_mesh(256, 256), S.hello_texture(320, 180), S.synthetic_grid(384, 216, n=113, layers=4), S.procedural(160, 90),
            S.features(), S.frag_depth(), S.multi_draw(), S.huge_triangles(), S.random_triangles(seed=11), S.random_lines(), S.random_points(),
            S.early_depth("force", "less"), S.early_depth("allow", "less-equal"), S.multiple_targets()]
    out += [S.fuzz(seed) for seed in range(0, 64, 4)]
    return out


def digest(frame) -> dict:
    import numpy as np
    d = {"color": hashlib.sha256(np.ascontiguousarray(frame.color).tobytes()).hexdigest()}
    if frame.depth is not None:
        d["depth"] = hashlib.sha256(np.ascontiguousarray(frame.depth).view(np.uint32).tobytes()).hexdigest()
    for k, extra in enumerate(frame.extra_colors or [], start=1):
        d[f"color{k}"] = hashlib.sha256(np.ascontiguousarray(extra).tobytes()).hexdigest()
    return d


def generate() -> dict:
    from oracle import pyoracle
    return {s.name: digest(pyoracle.render(s, want_coverage=False)) for s in scenes()}


if __name__ == "__main__":
    path = os.path.join(ROOT, "tests", "golden", "oracle_frames_sha256.json")
    with open(path, "w") as f:
        json.dump(generate(), f, indent=1, sort_keys=True)
        f.write("\n")
    print(path)
