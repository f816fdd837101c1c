"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_DIR, "liboracle.so")

MAX_GROUPS, MAX_BINDINGS, MAX_VB, MAX_ATTRS, MAX_COLOR = 4, 4, 8, 16, 4

TOPOLOGY = {"point-list": 0, "line-list": 1, "line-strip": 2, "triangle-list": 3, "triangle-strip": 4}
INDEX_FORMAT = {None: 0, "uint16": 1, "uint32": 2}
FRONT_FACE = {"ccw": 0, "cw": 1}
CULL_MODE = {None: 0, "front": 1, "back": 2}
COMPARE = {"never": 1, "less": 2, "equal": 3, "less-equal": 4, "greater": 5, "not-equal": 6,
           "greater-equal": 7, "always": 8}
FORMAT = {"rgba8unorm": 0, "rgba8unorm-srgb": 1, "bgra8unorm": 2, "bgra8unorm-srgb": 3, "r8unorm": 4,
          "rg8unorm": 5, "rgba8snorm": 6, "depth32float": 7}
ADDRESS = {"clamp-to-edge": 0, "repeat": 1, "mirror-repeat": 2}
STEP = {"vertex": 0, "instance": 1}
SHADER = {"colored_triangle": 0, "hello_shader": 0, "hello_mesh": 1, "hello_texture": 2, "procedural": 3,
          "features": 4, "frag_depth": 5, "early_force": 6, "early_allow": 7, "mrt": 8, "depth_only": 9, "prim_index": 10, "perspective": 11}
ATTR_SIZE = {"float32": 4, "float32x2": 8, "float32x3": 12, "float32x4": 16, "uint32": 4, "sint32": 4}


class Texture(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("format", C.c_uint32)]


class Sampler(C.Structure):
    _fields_ = [("address_u", C.c_uint32), ("address_v", C.c_uint32)]


class Buffer(C.Structure):
    _fields_ = [("data", C.c_void_p), ("size", C.c_uint64)]


class Binding(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("buffer", Buffer), ("texture", Texture), ("sampler", Sampler)]


class Bindings(C.Structure):
    _fields_ = [("b", (Binding * MAX_BINDINGS) * MAX_GROUPS)]


class VertexAttr(C.Structure):
    _fields_ = [("location", C.c_uint32), ("buffer", C.c_uint32), ("offset", C.c_uint32), ("size", C.c_uint32)]


class VbLayout(C.Structure):
    _fields_ = [("stride", C.c_uint32), ("step_mode", C.c_uint32)]


class Pipeline(C.Structure):
    _fields_ = [("shader", C.c_uint32), ("topology", C.c_uint32), ("strip_index_format", C.c_uint32),
                ("front_face", C.c_uint32), ("cull_mode", C.c_uint32), ("has_depth_state", C.c_uint32),
                ("depth_compare", C.c_uint32), ("depth_write", C.c_uint32), ("has_fragment", C.c_uint32),
                ("num_vertex_buffers", C.c_uint32), ("vb", VbLayout * MAX_VB),
                ("num_attrs", C.c_uint32), ("attrs", VertexAttr * MAX_ATTRS)]


class Pass(C.Structure):
    _fields_ = [("num_color", C.c_uint32), ("color", Texture * MAX_COLOR), ("color_clear", C.c_uint32 * MAX_COLOR),
                ("clear_color", (C.c_double * 4) * MAX_COLOR), ("has_depth", C.c_uint32), ("depth", Texture),
                ("depth_clear", C.c_uint32), ("clear_depth", C.c_float), ("ext_features", C.c_uint32)]


class RasterState(C.Structure):
    _fields_ = [("vp_x", C.c_float), ("vp_y", C.c_float), ("vp_w", C.c_float), ("vp_h", C.c_float),
                ("vp_min_depth", C.c_float), ("vp_max_depth", C.c_float),
                ("sc_x", C.c_uint32), ("sc_y", C.c_uint32), ("sc_w", C.c_uint32), ("sc_h", C.c_uint32),
                ("ext_features", C.c_uint32), ("color_write_mask", C.c_uint32 * MAX_COLOR),
                ("blend", (C.c_uint32 * 7) * MAX_COLOR), ("blend_constant", C.c_float * 4)]


class DrawDesc(C.Structure):
    _fields_ = [("indexed", C.c_uint32), ("first", C.c_uint32), ("count", C.c_uint32), ("base_vertex", C.c_int32),
                ("first_instance", C.c_uint32), ("instance_count", C.c_uint32), ("index_format", C.c_uint32),
                ("index_buffer", Buffer), ("vertex_buffers", Buffer * MAX_VB)]


class Stats(C.Structure):
    _fields_ = [("vertices_processed", C.c_uint64), ("primitives_assembled", C.c_uint64),
                ("primitives_culled", C.c_uint64), ("primitives_drawn", C.c_uint64),
                ("fragments_shaded", C.c_uint64), ("fragments_written", C.c_uint64)]


_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with the committed Makefile (gcc, -ffp-contract=off)."""
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_DIR, f)) > os.path.getmtime(_LIB_PATH)
            for f in ("oracle.cpp", "oracle_shaders.cpp", "oracle.h", "oracle_internal.h")):
        subprocess.check_call(["make", "-C", _DIR, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_texel_coordinate.restype = C.c_uint32
        _lib.orc_texel_coordinate.argtypes = [C.c_float, C.c_uint32, C.c_uint32]
        _lib.orc_f32_to_u8.restype = C.c_uint8
        _lib.orc_f32_to_u8.argtypes = [C.c_float]
        _lib.orc_texture_byte_size.restype = C.c_uint64
        _lib.orc_texture_byte_size.argtypes = [C.c_uint32] * 4
        _lib.orc_texel_byte_offset.restype = C.c_uint64
        _lib.orc_texel_byte_offset.argtypes = [C.c_uint32] * 6
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


@dataclass
class Frame:
    color: np.ndarray          # H x W x bpp uint8
    depth: np.ndarray | None   # H x W float32
    coverage: np.ndarray | None  # H x W uint32: fragments that reached the fragment stage
    stats: dict
    extra_colors: list = None  # further colour attachments, in location order


def render(scene, want_coverage: bool = True) -> Frame:
    """Render a wgpu_cpu_b200.scenes.Scene with the CPU oracle: State::load then every draw."""
    L = lib()
    keep = []  # keep numpy buffers alive
    W, H = scene.width, scene.height
    fmt = FORMAT[scene.color_format]
    bpp = {4: 1, 5: 2}.get(fmt, 4)
    color = (np.zeros((H, W, bpp), dtype=np.uint8) if scene.initial_color is None
             else np.ascontiguousarray(scene.initial_color, dtype=np.uint8).copy())
    depth = None
    if scene.has_depth:
        depth = (np.zeros((H, W), dtype=np.float32) if scene.initial_depth is None
                 else np.ascontiguousarray(scene.initial_depth, dtype=np.float32).copy())

    p = Pass()
    p.num_color = 1
    p.color[0] = Texture(_ptr(color), W, H, fmt)
    p.color_clear[0] = 1 if scene.clear_color is not None else 0
    if scene.clear_color is not None:
        for k in range(4):
            p.clear_color[0][k] = float(scene.clear_color[k])
    extra_colors = []
    for k, (xfmt, xclear) in enumerate(getattr(scene, "extra_targets", None) or [], start=1):
        xf = FORMAT[xfmt]
        arr = np.zeros((H, W, {4: 1, 5: 2}.get(xf, 4)), dtype=np.uint8)
        extra_colors.append(arr)
        p.color[k] = Texture(_ptr(arr), W, H, xf)
        p.color_clear[k] = 1 if xclear is not None else 0
        if xclear is not None:
            for c in range(4):
                p.clear_color[k][c] = float(xclear[c])
        p.num_color = k + 1
    p.ext_features = getattr(scene, "features", 0) & 23
    p.has_depth = 1 if scene.has_depth else 0
    if scene.has_depth:
        p.depth = Texture(_ptr(depth), W, H, FORMAT["depth32float"])
        p.depth_clear = 1 if scene.clear_depth is not None else 0
        p.clear_depth = float(scene.clear_depth) if scene.clear_depth is not None else 0.0

    pipe = Pipeline()
    pipe.shader = SHADER[scene.shader]
    pipe.topology = TOPOLOGY[scene.topology]
    pipe.strip_index_format = INDEX_FORMAT[scene.strip_index_format]
    pipe.front_face = FRONT_FACE[scene.front_face]
    pipe.cull_mode = CULL_MODE[scene.cull_mode]
    pipe.has_depth_state = 1 if scene.depth_compare is not None else 0
    pipe.depth_compare = COMPARE[scene.depth_compare] if scene.depth_compare is not None else 0
    pipe.depth_write = 1 if scene.depth_write else 0
    pipe.has_fragment = 1
    pipe.num_vertex_buffers = len(scene.vertex_layouts)
    na = 0
    for bi, lay in enumerate(scene.vertex_layouts):
        pipe.vb[bi] = VbLayout(lay.stride, STEP[lay.step_mode])
        for a in lay.attributes:
            pipe.attrs[na] = VertexAttr(a.location, bi, a.offset, ATTR_SIZE[a.format])
            na += 1
    pipe.num_attrs = na

    rs = RasterState()
    L.orc_default_raster_state(W, H, C.byref(rs))
    if scene.viewport is not None:
        rs.vp_x, rs.vp_y, rs.vp_w, rs.vp_h, rs.vp_min_depth, rs.vp_max_depth = scene.viewport
    if scene.scissor is not None:
        rs.sc_x, rs.sc_y, rs.sc_w, rs.sc_h = scene.scissor
    rs.ext_features = getattr(scene, "features", 0) & 23       # depth range, write mask, sRGB encode, blend (oracle.h ORC_EXT_*)
    if getattr(scene, "blend", None):
        BLEND_FACTOR = {"zero": 0, "one": 1, "src": 2, "one-minus-src": 3, "src-alpha": 4, "one-minus-src-alpha": 5, "dst": 6,
                        "one-minus-dst": 7, "dst-alpha": 8, "one-minus-dst-alpha": 9, "src-alpha-saturated": 10, "constant": 11,
                        "one-minus-constant": 12}           # wgpu::BlendFactor, numbering of include/wgpu_b200.h
        BLEND_OPERATION = {"add": 0, "subtract": 1, "reverse-subtract": 2, "min": 3, "max": 4}
        (cs, cd, co), (as_, ad, ao) = scene.blend["color"], scene.blend["alpha"]
        for k, v in enumerate([1, BLEND_FACTOR[cs], BLEND_FACTOR[cd], BLEND_OPERATION[co], BLEND_FACTOR[as_], BLEND_FACTOR[ad], BLEND_OPERATION[ao]]):
            rs.blend[0][k] = v
        for k in range(4):
            rs.blend_constant[k] = float(scene.blend_constant[k])
    rs.color_write_mask[0] = getattr(scene, "color_write_mask", 15)
    for k in range(1, 4):
        rs.color_write_mask[k] = 15
    dyn = {}
    if getattr(scene, "features", 0) & 8 and scene.dynamic_offsets:       # dynamic offsets: the k-th offset moves the k-th dynamic binding
        for g, offs in scene.dynamic_offsets.items():
            for b, off in zip(sorted(scene.dynamic_bindings[g]), offs):
                dyn[(g, b)] = off

    binds = Bindings()
    for (g, b), res in scene.bindings.items():
        e = binds.b[g][b]
        if res[0] == "buffer":
            arr = np.ascontiguousarray(res[1]).view(np.uint8).reshape(-1)
            arr = arr[dyn.get((g, b), 0):]
            keep.append(arr)
            e.kind = 1
            e.buffer = Buffer(_ptr(arr), arr.nbytes)
        elif res[0] == "texture":
            arr = np.ascontiguousarray(res[1], dtype=np.uint8)
            keep.append(arr)
            e.kind = 2
            e.texture = Texture(_ptr(arr), arr.shape[1], arr.shape[0], FORMAT[res[2]])
        elif res[0] == "sampler":
            e.kind = 3
            e.sampler = Sampler(ADDRESS[res[1]], ADDRESS[res[2]])

    coverage = np.zeros((H, W), dtype=np.uint32) if want_coverage else None
    stats = Stats()

    e = L.orc_pass_load(C.byref(p))
    if e:
        raise RuntimeError(f"orc_pass_load failed: {e}")
    vbs = [np.ascontiguousarray(vb).view(np.uint8).reshape(-1) for vb in scene.vertex_buffers]
    idx = None
    if scene.index_data is not None:
        idx = np.ascontiguousarray(scene.index_data)
    for d in scene.draws:
        dd = DrawDesc()
        dd.indexed = 1 if d.indexed else 0
        dd.first, dd.count, dd.base_vertex = d.first, d.count, d.base_vertex
        dd.first_instance, dd.instance_count = d.first_instance, d.instance_count
        if d.indexed:
            dd.index_format = 1 if idx.dtype == np.uint16 else 2
            dd.index_buffer = Buffer(_ptr(idx), idx.nbytes)
        for bi, vb in enumerate(vbs):
            dd.vertex_buffers[bi] = Buffer(_ptr(vb), vb.nbytes)
        e = L.orc_draw_execute(C.byref(p), C.byref(pipe), C.byref(rs), C.byref(binds), C.byref(dd), C.byref(stats),
                               _ptr(coverage) if coverage is not None else None)
        if e:
            raise RuntimeError(f"orc_draw_execute failed: {e}")
    return Frame(color, depth, coverage, {f[0]: int(getattr(stats, f[0])) for f in Stats._fields_}, extra_colors)
