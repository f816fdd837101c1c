#!/usr/bin/env python
"""More fuzz seeds than the suite carries, on the software model of tests/cusim (no GPU needed):
    python tools/fuzz_on_model.py <first seed> <last seed> [generator: fuzz|fuzz_shaders|fuzz_features|fuzz_textured|hiz]
Every scene is rendered with and without the coverage capture (= without and with the hierarchical depth test) and
compared with the oracle bit for bit.  `hiz` is a generator for the depth bound: layered jittered meshes like C3 at
random sizes, layer counts, jitter and draw order, Less / LessEqual + depth write."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("WGB_CUSIM", "1")
os.environ.setdefault("CUSIM_THREADS", "4")


def hiz_scene(seed):
    from wgpu_cpu_b200 import scenes as S
    rng = np.random.default_rng(seed)
    w, h = int(rng.integers(40, 400)), int(rng.integers(40, 300))
    n, layers = int(rng.integers(8, 90)), int(rng.integers(1, 6))
    s = S.synthetic_grid(w, h, n=n, layers=layers, config_id=1000 + seed)
    v = s.vertex_buffers[0].view(np.float32).reshape(-1, 8).copy()
    # stronger / weaker depth jitter, layers that interleave, a few vertices pushed outside the clip volume
    v[:, 2] = np.clip(v[:, 2] + rng.normal(0.0, rng.choice([0.0, 0.02, 0.1, 0.3]), len(v)).astype(np.float32), 0.0, 1.0)
    if rng.random() < 0.3:
        v[:, 0] *= np.float32(rng.uniform(0.8, 1.3))
    if rng.random() < 0.5:          # back to front, or shuffled layers
        idx = s.index_data.reshape(layers, -1)
        s.index_data = np.ascontiguousarray(idx[rng.permutation(layers)].reshape(-1))
    s.vertex_buffers = [v.view(np.uint8).reshape(-1)]
    s.depth_compare = str(rng.choice(["less", "less-equal"]))
    s.name = f"hiz_{seed}"
    return s


def main():
    first, last = int(sys.argv[1]), int(sys.argv[2])
    gen = sys.argv[3] if len(sys.argv) > 3 else "fuzz"
    from tests.cusim import build as cusim_build
    from wgpu_cpu_b200 import api, scenes as S
    api.LIB_PATH = cusim_build.build()
    from oracle import pyoracle
    from wgpu_cpu_b200.render import render_scene
    devices = {}

    def device_for(features):          # the opt-in features are a property of the device (wgb_device_descriptor.features)
        if features not in devices:
            devices[features] = api.instance().request_adapter().request_device(0, features=features)
        return devices[features]

    make = hiz_scene if gen == "hiz" else getattr(S, gen)
    bad = 0
    for seed in range(first, last + 1):
        sc = make(seed)
        dev, queue = device_for(int(getattr(sc, "features", 0) or 0))
        ref = pyoracle.render(sc)
        for cov in ((False,) if gen == "fuzz_features" else (True, False)):
            got = render_scene(dev, queue, sc, want_coverage=cov)
            ok = np.array_equal(got.color, ref.color) and (ref.depth is None or np.array_equal(got.depth.view(np.uint32), ref.depth.view(np.uint32)))
            if cov:
                ok = ok and np.array_equal(got.coverage, ref.coverage)
            if not ok:
                bad += 1
                print(f"MISMATCH {gen} seed {seed} coverage_capture={cov} ({sc.name})", flush=True)
        if (seed - first) % 50 == 49:
            print(f"{gen}: seeds {first}..{seed} done, {bad} mismatches", flush=True)
    print(f"{gen}: seeds {first}..{last}: {bad} mismatches")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
