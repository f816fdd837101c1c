"""Pack the reference's example meshes into small binary fixtures.

The GPU box has no /root/reference, so the two public test models the BASELINE
configs name (Utah teapot for C1, Stanford bunny for C2/C5;
wgpu-cpu/examples/teapot.obj, stanford-bunny.obj) are packed once, here, into
`meshes.npz` (float32 positions, uint32 triangle indices -- what tobj's
GPU_LOAD_OPTIONS hands hello_mesh.rs:536-538).  Run in the build container:

    python tests/assets/make_assets.py
"""
import os
import sys

import numpy as np

REF = "/root/reference/wgpu-cpu/examples"


def load_obj(path):
    pos, idx = [], []
    with open(path) as f:
        for line in f:
            if line.startswith("v "):
                pos.append([float(t) for t in line.split()[1:4]])
            elif line.startswith("f "):
                # position index is the first field of each a/b/c triple; fan-triangulate
                ids = [int(t.split("/")[0]) for t in line.split()[1:]]
                ids = [i - 1 if i > 0 else len(pos) + i for i in ids]
                for k in range(1, len(ids) - 1):
                    idx.append([ids[0], ids[k], ids[k + 1]])
    return np.asarray(pos, dtype=np.float32), np.asarray(idx, dtype=np.uint32).reshape(-1)


def main():
    out = {}
    for name, fn in (("teapot", "teapot.obj"), ("bunny", "stanford-bunny.obj"), ("cube", "cube.obj")):
        p, i = load_obj(os.path.join(REF, fn))
        print(name, p.shape, i.shape, file=sys.stderr)
        out[name + "_positions"] = p
        out[name + "_indices"] = i
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "meshes.npz"), **out)


if __name__ == "__main__":
    main()
