"""Regenerate tests/golden/oracle_frames.json: digests of the CPU oracle's output for the parity scenes.

The reference's own golden images (tests/reference/*.png) are git-LFS pointer files, so no reference
pixels exist to pin against; these digests pin the ORACLE against accidental change (any edit to
oracle/*.cpp that alters a frame shows up here) and record the per-scene counters.  Run:

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pyoracle  # noqa: E402
from wgpu_cpu_b200 import scenes as S  # noqa: E402


def scene_list():
    out = [S.colored_triangle(v) for v in ("default", "cull_front", "draw_backwards", "draw_backwards_no_cull", "lines")]
    out += [S.hello_mesh(512, 512), S.hello_texture(640, 360), S.synthetic_grid(384, 216, n=113, layers=4),
            S.procedural(480, 270), S.random_triangles(seed=11), S.quad_strip(), S.random_lines(), S.random_points(),
            S.features(), S.frag_depth()]
    return out


def digest(frame):
    h = lambda a: hashlib.sha256(a.tobytes()).hexdigest()[:32]
    return {"color": h(frame.color), "depth": h(frame.depth) if frame.depth is not None else None,
            "coverage": h(frame.coverage), "stats": frame.stats}


def main():
    out = {s.name: digest(pyoracle.render(s)) for s in scene_list()}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_frames.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(f"wrote {len(out)} scene digests")


if __name__ == "__main__":
    main()
