"""Shader registry: WGSL sources of the BASELINE configs / parity tests and, next to them, the
CUDA C++ the WGSL->CUDA emitter produces for their entry points (`shaders/emitted/*.cuh`, kept as
golden emitter output)."""
from __future__ import annotations

import os

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shaders")
NAMES = ("index_triangle", "mesh_vertex_color", "mesh_textured", "procedural", "features", "frag_depth", "early_force", "early_allow", "mrt", "depth_only", "prim_index", "perspective")
# the scenes keep the names of the reference programs they restate (scenes.py, oracle shader ids)
ALIASES = {"colored_triangle": "index_triangle", "hello_shader": "index_triangle", "hello_mesh": "mesh_vertex_color",
           "hello_texture": "mesh_textured"}


def wgsl(name: str) -> str:
    name = ALIASES.get(name, name)
    with open(os.path.join(_DIR, name + ".wgsl")) as f:
        return f.read()


def emitted(name: str, stage: str) -> str:
    """Golden emitter output for `vs_main` ("vs") / `fs_main` ("fs") of shader `name`."""
    name = ALIASES.get(name, name)
    with open(os.path.join(_DIR, "emitted", f"{name}.{stage}.cuh")) as f:
        return f.read()
