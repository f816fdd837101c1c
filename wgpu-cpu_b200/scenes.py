"""Scene descriptions for the BASELINE configs and the parity tests.

A `Scene` is a neutral, backend-independent description of one render pass: the
same object is rendered by the CPU oracle (tests only) and by the B200 backend
through the wgpu-style host API (`wgpu_cpu_b200.api`), so both consume
byte-identical inputs.

Scene definitions follow SURVEY.md section 8(d):
  C1 hello_mesh    -- wgpu-cpu/examples/hello_mesh.rs:144-157,534-682
  C2 hello_texture -- wgpu-cpu/examples/hello_texture.rs:235-240,647-677
  C3 synthetic     -- jittered grid, splitmix64
  C4 procedural    -- full-screen 64-iteration fragment shader
  colored_triangle -- wgpu-cpu-tests/src/tests/colored_triangle.rs:140-198
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ASSETS = os.path.join(os.path.dirname(_HERE), "tests", "assets", "meshes.npz")


@dataclass
class VertexAttribute:
    location: int
    format: str  # "float32", "float32x2", "float32x3", "float32x4", "uint32"
    offset: int


@dataclass
class VertexBufferLayout:
    stride: int
    step_mode: str  # "vertex" | "instance"
    attributes: List[VertexAttribute]


@dataclass
class Draw:
    indexed: bool
    first: int
    count: int
    base_vertex: int = 0
    first_instance: int = 0
    instance_count: int = 1


@dataclass
class Scene:
    name: str
    width: int
    height: int
    shader: str  # key into wgpu_cpu_b200.shaders.SHADERS
    color_format: str = "rgba8unorm-srgb"
    has_depth: bool = True
    clear_color: Optional[Tuple[float, float, float, float]] = (0.0, 0.0, 0.0, 1.0)  # None = LoadOp::Load
    clear_depth: Optional[float] = 1.0
    topology: str = "triangle-list"
    strip_index_format: Optional[str] = None
    front_face: str = "ccw"
    cull_mode: Optional[str] = None
    depth_compare: Optional[str] = "less"  # None = pipeline without depth-stencil state
    depth_write: bool = True
    vertex_layouts: List[VertexBufferLayout] = field(default_factory=list)
    vertex_buffers: List[np.ndarray] = field(default_factory=list)  # uint8 arrays
    index_data: Optional[np.ndarray] = None  # uint16 / uint32
    bindings: Dict[Tuple[int, int], tuple] = field(default_factory=dict)
    # ("buffer", uint8 array) | ("texture", HxWx4 uint8, format) | ("sampler", addr_u, addr_v)
    draws: List[Draw] = field(default_factory=list)
    viewport: Optional[Tuple[float, float, float, float, float, float]] = None
    scissor: Optional[Tuple[int, int, int, int]] = None
    # opt-in behaviour beyond the reference (api.FEATURE bits; the device must be requested with them)
    features: int = 0
    color_write_mask: int = 15                       # api.COLOR_WRITE bits, honoured with FEATURE["COLOR_WRITE_MASK"]
    dynamic_offsets: Optional[dict] = None           # {group: [offsets]} for bindings listed in dynamic_bindings
    dynamic_bindings: Optional[dict] = None          # {group: [binding numbers with has_dynamic_offset]}
    blend: Optional[dict] = None                     # {"color": (src, dst, op), "alpha": (src, dst, op)}, FEATURE["BLEND"]
    blend_constant: Tuple[float, float, float, float] = (0.0, 0.0, 0.0, 0.0)
    initial_color: Optional[np.ndarray] = None  # for LoadOp::Load passes
    extra_targets: Optional[list] = None        # further colour attachments (@location(1..)): [(format, clear colour | None)]
    initial_depth: Optional[np.ndarray] = None

    @property
    def num_primitives(self) -> int:
        n = 0
        for d in self.draws:
            if self.topology == "triangle-list":
                n += (d.count // 3) * d.instance_count
            elif self.topology == "triangle-strip":
                n += max(d.count - 2, 0) * d.instance_count
            elif self.topology == "line-list":
                n += (d.count // 2) * d.instance_count
            elif self.topology == "line-strip":
                n += max(d.count - 1, 0) * d.instance_count
            else:
                n += d.count * d.instance_count
        return n

    def algorithmic_bytes(self) -> int:
        """4*I (2*I for u16) + stride*V_unique + U + (Bc+Bd)*W*H + Tex."""
        b = 0
        if self.index_data is not None:
            b += self.index_data.nbytes
        for vb in self.vertex_buffers:
            b += vb.nbytes
        for res in self.bindings.values():
            if res[0] == "buffer":
                b += res[1].nbytes
            elif res[0] == "texture":
                b += res[1].nbytes
        b += 4 * self.width * self.height
        if self.has_depth:
            b += 4 * self.width * self.height
        return b


# --------------------------------------------------------------------------- #
# shared RNG: splitmix64, state0 = 0x9E3779B97F4A7C15 ^ config_id
# --------------------------------------------------------------------------- #
_GOLDEN = np.uint64(0x9E3779B97F4A7C15)


def splitmix64_u01(n: int, config_id: int) -> np.ndarray:
    """n values of u01 = (next() >> 40) * 2^-24 from splitmix64 (vectorised)."""
    with np.errstate(over="ignore"):
        state0 = np.uint64(0x9E3779B97F4A7C15 ^ config_id)
        z = state0 + _GOLDEN * np.arange(1, n + 1, dtype=np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


# --------------------------------------------------------------------------- #
# camera math (nalgebra Perspective3 / Isometry3::face_towards, f32)
# --------------------------------------------------------------------------- #
def _normalize(v):
    v = np.asarray(v, dtype=np.float32)
    return (v / np.float32(np.sqrt(np.float32(np.dot(v, v))))).astype(np.float32)


def camera_matrix(eye, target, up, aspect, fovy=math.pi / 4, znear=0.001, zfar=100.0) -> np.ndarray:
    """CameraBufferData::new (hello_mesh.rs:671-681): P' * inverse(face_towards(eye, target, up)),
    P = Perspective3(aspect, fovy, near, far) with P[2][2] *= -1, P[3][2] = 1.  Returns the
    64 uniform bytes (column-major mat4x4f)."""
    f32 = np.float32
    eye = np.asarray(eye, dtype=f32)
    target = np.asarray(target, dtype=f32)
    zaxis = _normalize(target - eye)
    xaxis = _normalize(np.cross(np.asarray(up, dtype=f32), zaxis))
    yaxis = np.cross(zaxis, xaxis).astype(f32)
    rot = np.stack([xaxis, yaxis, zaxis], axis=1).astype(f32)  # columns
    view = np.eye(4, dtype=f32)
    view[:3, :3] = rot.T
    view[:3, 3] = -(rot.T @ eye)
    proj = np.zeros((4, 4), dtype=f32)
    t = f32(math.tan(fovy / 2.0))
    proj[0, 0] = f32(1.0) / (f32(aspect) * t)
    proj[1, 1] = f32(1.0) / t
    proj[2, 2] = f32(zfar + znear) / f32(znear - zfar)
    proj[2, 3] = f32(2.0 * zfar * znear) / f32(znear - zfar)
    proj[3, 2] = f32(-1.0)
    proj[2, 2] *= f32(-1.0)
    proj[3, 2] = f32(1.0)
    m = (proj @ view).astype(f32)
    return np.ascontiguousarray(m.T).view(np.uint8).reshape(-1).copy()  # column-major bytes


def identity_matrix_bytes() -> np.ndarray:
    return np.eye(4, dtype=np.float32).view(np.uint8).reshape(-1).copy()


def _rotation_y(angle: float) -> np.ndarray:
    c, s = np.float32(math.cos(angle)), np.float32(math.sin(angle))
    r = np.eye(3, dtype=np.float32)
    r[0, 0], r[0, 2], r[2, 0], r[2, 2] = c, s, -s, c
    return r


# --------------------------------------------------------------------------- #
# meshes
# --------------------------------------------------------------------------- #
def load_mesh(name: str):
    data = np.load(_ASSETS)
    return data[name + "_positions"].astype(np.float32), data[name + "_indices"].astype(np.uint32)


def _bounds(pos):
    mn = pos.min(axis=0).astype(np.float32)
    mx = pos.max(axis=0).astype(np.float32)
    size = (mx - mn).astype(np.float32)
    center = (np.float32(0.5) * (mn + mx)).astype(np.float32)
    return mn, size, center


_POS_COLOR_LAYOUT = VertexBufferLayout(32, "vertex", [VertexAttribute(0, "float32x4", 0), VertexAttribute(1, "float32x4", 16)])
_POS_UV_LAYOUT = VertexBufferLayout(24, "vertex", [VertexAttribute(0, "float32x4", 0), VertexAttribute(1, "float32x2", 16)])


def hello_mesh(width=512, height=512, mesh="teapot") -> Scene:
    """C1 (hello_mesh.rs): vertex = {pos vec4 (w=1), colour = (pos-min)/size, a=1}, u32 indices,
    TriangleList, front=Cw, cull=Back, Depth32Float Less+write, clear black / 1.0."""
    pos, idx = load_mesh(mesh)
    mn, size, center = _bounds(pos)
    v = np.ones((pos.shape[0], 8), dtype=np.float32)
    v[:, 0:3] = pos
    v[:, 4:7] = (pos - mn) / size
    eye = center + np.array([0.0, 0.0, -float(size.max())], dtype=np.float32)
    target = center + np.array([0.0, float(size[1]) * 0.25, 0.0], dtype=np.float32)
    cam = camera_matrix(eye, target, (0.0, 1.0, 0.0), width / height)
    return Scene(
        name=f"hello_mesh_{mesh}_{width}x{height}", width=width, height=height, shader="hello_mesh",
        front_face="cw", cull_mode="back",
        vertex_layouts=[_POS_COLOR_LAYOUT], vertex_buffers=[v.view(np.uint8).reshape(-1)],
        index_data=idx, bindings={(0, 0): ("buffer", cam)},
        draws=[Draw(True, 0, int(idx.size))],
    )


def test_card(size=512) -> np.ndarray:
    """Procedural test card: 8x8 checker of size/8-px cells, cell colour
    (cx*36, cy*36, (cx^cy)*36, 255), plus a 1-px white grid on every cell boundary."""
    cell = size // 8
    yy, xx = np.meshgrid(np.arange(size), np.arange(size), indexing="ij")
    cx, cy = xx // cell, yy // cell
    img = np.zeros((size, size, 4), dtype=np.uint8)
    img[..., 0] = cx * 36
    img[..., 1] = cy * 36
    img[..., 2] = (cx ^ cy) * 36
    img[..., 3] = 255
    grid = (xx % cell == 0) | (yy % cell == 0)
    img[grid] = 255
    return img


def hello_texture(width=1920, height=1080, mesh="bunny", yaw=0.0) -> Scene:
    """C2 (hello_texture.rs): vertex = {pos vec4, uv vec2}, uv = (atan2(n.x,n.z)/tau + 0.5,
    0.5*n.z + 0.5)*5 with n = normalize(pos-center); eye = center + (0,0,1.5*max(size));
    sampler Repeat/Nearest over the procedural test card; pipeline state as C1."""
    pos, idx = load_mesh(mesh)
    mn, size, center = _bounds(pos)
    d = (pos - center).astype(np.float32)
    n = d / np.sqrt((d * d).sum(axis=1, keepdims=True)).astype(np.float32)
    uv = np.stack([np.arctan2(n[:, 0], n[:, 2]) / np.float32(2 * math.pi) + np.float32(0.5),
                   np.float32(0.5) * n[:, 2] + np.float32(0.5)], axis=1).astype(np.float32) * np.float32(5.0)
    v = np.ones((pos.shape[0], 6), dtype=np.float32)
    v[:, 0:3] = pos
    v[:, 4:6] = uv
    offset = _rotation_y(yaw) @ np.array([0.0, 0.0, 1.5 * float(size.max())], dtype=np.float32)
    eye = center + offset.astype(np.float32)
    cam = camera_matrix(eye, center, (0.0, 1.0, 0.0), width / height)
    return Scene(
        name=f"hello_texture_{mesh}_{width}x{height}", width=width, height=height, shader="hello_texture",
        front_face="cw", cull_mode="back",
        vertex_layouts=[_POS_UV_LAYOUT], vertex_buffers=[v.view(np.uint8).reshape(-1)],
        index_data=idx,
        bindings={(0, 0): ("buffer", cam), (1, 0): ("texture", test_card(512), "rgba8unorm-srgb"),
                  (1, 1): ("sampler", "repeat", "repeat")},
        draws=[Draw(True, 0, int(idx.size))],
    )


def synthetic_grid(width=3840, height=2160, n=1119, layers=4, config_id=3) -> Scene:
    """C3: `layers` jittered n x n vertex grids covering NDC; vertex (l,i,j):
    x = -1 + 2(i + 0.35(2u-1))/(n-1), y likewise, z = (l + 0.5 + 0.4(2u-1))/layers, w = 1,
    colour = 3 x u01, a = 1; identity camera; cull none, Less+write.  n=1119, layers=4 gives
    9 999 392 triangles."""
    nv = layers * n * n
    u = splitmix64_u01(nv * 6, config_id).reshape(nv, 6)
    l, i, j = np.meshgrid(np.arange(layers, dtype=np.float32), np.arange(n, dtype=np.float32),
                          np.arange(n, dtype=np.float32), indexing="ij")
    f32 = np.float32
    v = np.ones((nv, 8), dtype=np.float32)
    v[:, 0] = (f32(-1.0) + f32(2.0) * (i.reshape(-1) + f32(0.35) * (f32(2.0) * u[:, 0] - f32(1.0))) / f32(n - 1))
    v[:, 1] = (f32(-1.0) + f32(2.0) * (j.reshape(-1) + f32(0.35) * (f32(2.0) * u[:, 1] - f32(1.0))) / f32(n - 1))
    v[:, 2] = (l.reshape(-1) + f32(0.5) + f32(0.4) * (f32(2.0) * u[:, 2] - f32(1.0))) / f32(layers)
    v[:, 4:7] = u[:, 3:6]
    # indices: per layer, per cell two triangles
    ci, cj = np.meshgrid(np.arange(n - 1, dtype=np.uint32), np.arange(n - 1, dtype=np.uint32), indexing="ij")
    v00 = (ci * n + cj).reshape(-1)
    v01 = v00 + 1
    v10 = v00 + n
    v11 = v10 + 1
    cell = np.stack([v00, v10, v01, v01, v10, v11], axis=1).astype(np.uint32)
    idx = np.concatenate([cell + np.uint32(k * n * n) for k in range(layers)], axis=0).reshape(-1)
    return Scene(
        name=f"synthetic_{layers}x{n}x{n}_{width}x{height}", width=width, height=height, shader="hello_mesh",
        front_face="ccw", cull_mode=None,
        vertex_layouts=[_POS_COLOR_LAYOUT], vertex_buffers=[v.view(np.uint8).reshape(-1)],
        index_data=idx, bindings={(0, 0): ("buffer", identity_matrix_bytes())},
        draws=[Draw(True, 0, int(idx.size))],
    )


def procedural(width=7680, height=4320) -> Scene:
    """C4: 2 triangles from vertex_index, no vertex buffer, no depth attachment, Rgba8Unorm."""
    return Scene(
        name=f"procedural_{width}x{height}", width=width, height=height, shader="procedural",
        color_format="rgba8unorm", has_depth=False, clear_depth=None, depth_compare=None, depth_write=False,
        front_face="ccw", cull_mode=None, draws=[Draw(False, 0, 6)],
    )


def colored_triangle(variant="default", width=512, height=512) -> Scene:
    """The five image-regression scenes of wgpu-cpu-tests (colored_triangle.rs:140-198):
    512x512 Rgba8UnormSrgb + Depth32Float, clear black / 1.0, Less."""
    s = Scene(name=f"colored_triangle_{variant}", width=width, height=height, shader="colored_triangle",
              draws=[Draw(False, 0, 3)])
    if variant == "default":
        s.front_face, s.cull_mode = "cw", "back"
    elif variant == "cull_front":
        s.front_face, s.cull_mode = "cw", "front"
    elif variant == "draw_backwards":
        s.front_face, s.cull_mode = "ccw", "back"
    elif variant == "draw_backwards_no_cull":
        s.front_face, s.cull_mode = "ccw", None
    elif variant == "lines":
        s.front_face, s.cull_mode = "ccw", None
        s.topology = "line-strip"
        s.draws = [Draw(False, 0, 4)]
    else:
        raise ValueError(variant)
    return s


# --------------------------------------------------------------------------- #
# extra parity scenes (edge cases the reference's own tests do not cover end to end)
# --------------------------------------------------------------------------- #
def random_triangles(width=256, height=192, count=400, seed=11, spread=1.6, with_w=True, **overrides) -> Scene:
    """Random triangles, many of them crossing the clip volume (all six planes) and with
    varying w, to exercise the clipper, the clipped-depth quirk and order-dependent
    double coverage of shared edges."""
    u = splitmix64_u01(count * 3 * 8, seed).reshape(count * 3, 8)
    v = np.ones((count * 3, 8), dtype=np.float32)
    v[:, 0] = (u[:, 0] * 2 - 1) * np.float32(spread)
    v[:, 1] = (u[:, 1] * 2 - 1) * np.float32(spread)
    v[:, 2] = u[:, 2] * np.float32(1.4) - np.float32(0.2)
    if with_w:
        v[:, 3] = np.float32(0.25) + u[:, 3] * np.float32(1.5)
    # shrink each triangle around its first vertex so sizes vary
    tri = v.reshape(count, 3, 8)
    scale = (u.reshape(count, 3, 8)[:, 0, 7] ** 2).reshape(count, 1, 1).astype(np.float32)
    tri[:, 1:, 0:3] = tri[:, :1, 0:3] + (tri[:, 1:, 0:3] - tri[:, :1, 0:3]) * scale
    v = tri.reshape(count * 3, 8)
    v[:, 4:7] = u[:, 4:7]
    s = Scene(
        name=f"random_triangles_{count}_{seed}", width=width, height=height, shader="hello_mesh",
        vertex_layouts=[_POS_COLOR_LAYOUT], vertex_buffers=[np.ascontiguousarray(v).view(np.uint8).reshape(-1)],
        bindings={(0, 0): ("buffer", identity_matrix_bytes())}, draws=[Draw(False, 0, count * 3)],
    )
    for k, val in overrides.items():
        setattr(s, k, val)
    return s


def quad_strip(width=200, height=150, rows=6, cols=40, restart=True, index_format="uint16") -> Scene:
    """Triangle strips with primitive restart (index.rs:90-106, primitive.rs:407-487)."""
    f32 = np.float32
    vs = []
    for r in range(rows + 1):
        for c in range(cols + 1):
            x = f32(-0.95 + 1.9 * c / cols)
            y = f32(-0.9 + 1.8 * r / rows)
            vs.append([x, y, f32(0.1 + 0.8 * ((r * 7 + c * 3) % 11) / 11.0), 1.0,
                       c / cols, r / rows, ((r + c) % 2), 1.0])
    v = np.asarray(vs, dtype=np.float32)
    sep = 0xFFFF if index_format == "uint16" else 0xFFFFFFFF
    idx = []
    for r in range(rows):
        for c in range(cols + 1):
            idx += [r * (cols + 1) + c, (r + 1) * (cols + 1) + c]
        if restart:
            idx.append(sep)
            if r % 2 == 1:
                idx += [0, sep]  # a lone vertex between separators must be dropped
    idx = np.asarray(idx, dtype=np.uint16 if index_format == "uint16" else np.uint32)
    return Scene(
        name=f"quad_strip_{index_format}_{int(restart)}", width=width, height=height, shader="hello_mesh",
        topology="triangle-strip", strip_index_format=index_format if restart else None,
        vertex_layouts=[_POS_COLOR_LAYOUT], vertex_buffers=[v.view(np.uint8).reshape(-1)],
        index_data=idx, bindings={(0, 0): ("buffer", identity_matrix_bytes())},
        draws=[Draw(True, 0, int(idx.size))],
    )


def random_lines(width=160, height=120, count=60, seed=5, topology="line-list") -> Scene:
    u = splitmix64_u01(count * 2 * 8, seed).reshape(count * 2, 8)
    v = np.ones((count * 2, 8), dtype=np.float32)
    v[:, 0] = (u[:, 0] * 2 - 1) * np.float32(1.5)
    v[:, 1] = (u[:, 1] * 2 - 1) * np.float32(1.5)
    v[:, 2] = u[:, 2] * np.float32(1.3) - np.float32(0.15)
    v[:, 3] = np.float32(0.5) + u[:, 3]
    v[:, 4:7] = u[:, 4:7]
    return Scene(
        name=f"random_lines_{topology}_{count}", width=width, height=height, shader="hello_mesh", topology=topology,
        vertex_layouts=[_POS_COLOR_LAYOUT], vertex_buffers=[v.view(np.uint8).reshape(-1)],
        bindings={(0, 0): ("buffer", identity_matrix_bytes())}, draws=[Draw(False, 0, count * 2)],
    )


def random_points(width=96, height=64, count=500, seed=9) -> Scene:
    s = random_lines(width, height, count // 2, seed, "point-list")
    s.name = f"random_points_{count}"
    return s


def features(width=200, height=160, instances=3) -> Scene:
    """instance_index, a flat u32 varying, front_facing and discard (shaders/features.wgsl)."""
    base = random_triangles(width, height, count=60, seed=21, spread=0.9, with_w=False)
    params = np.zeros(20, dtype=np.float32)
    params[:16] = np.eye(4, dtype=np.float32).reshape(-1)
    params[16:20] = [0.15, -0.1, 0.05, 0.0]
    base.name = "features"
    base.shader = "features"
    base.bindings = {(0, 0): ("buffer", params.view(np.uint8).reshape(-1).copy())}
    base.draws = [Draw(False, 0, 180, 0, 1, instances)]
    return base


def frag_depth(width=160, height=120) -> Scene:
    s = random_triangles(width, height, count=80, seed=31, spread=1.1, with_w=False)
    s.name = "frag_depth"
    s.shader = "frag_depth"
    return s


def early_depth(kind="force", compare="less", width=160, height=120, topology="triangle-list") -> Scene:
    """@early_depth_test in front of a stage that discards and writes frag_depth (shaders/early_force.wgsl,
    early_allow.wgsl; fragment.rs:166-194)."""
    if topology == "triangle-list":
        s = random_triangles(width, height, count=80, seed=33, spread=1.1, with_w=False)
    else:
        s = random_lines(width, height, 60, 34, topology)
    s.name = f"early_{kind}_{compare}_{topology}"
    s.shader = f"early_{kind}"
    s.depth_compare = compare
    s.depth_write = True
    s.clear_depth = 0.5 if compare.startswith("greater") else 1.0
    return s


def multiple_targets(width=180, height=130, compare="less", second_load=False) -> Scene:
    """Three colour attachments of different texel sizes written by one fragment stage whose outputs are declared out
    of location order (shaders/mrt.wgsl; fragment.rs:457-488 visits them in declaration order and runs the late depth
    test once, at the first)."""
    s = random_triangles(width, height, count=120, seed=44, spread=1.1, with_w=False)
    s.name = f"multiple_targets_{compare}" + ("_load" if second_load else "")
    s.shader = "mrt"
    s.color_format = "rgba8unorm"
    s.depth_compare, s.depth_write = compare, True
    s.clear_depth = 0.5 if compare.startswith("greater") else 1.0
    s.extra_targets = [("bgra8unorm", None if second_load else (0.25, 0.5, 0.75, 1.0)), ("rg8unorm", (0.5, 0.125, 0.0, 0.0))]
    return s


def multi_draw(width=192, height=128) -> Scene:
    """Several draws in one pass over the same attachments (later draws see earlier depth/colour), u16 indices with a
    base_vertex, a ragged index count (the incomplete tail primitive is dropped, util/mod.rs:58-76) and an empty draw."""
    base = random_triangles(width, height, count=90, seed=61, spread=1.2, with_w=False)
    idx = np.arange(270, dtype=np.uint16)
    idx16 = np.concatenate([idx[:90], (idx[90:180] - 60).astype(np.uint16)])   # second range uses base_vertex = 60
    base.name = "multi_draw"
    base.index_data = idx16
    base.draws = [Draw(True, 0, 90), Draw(False, 180, 90), Draw(True, 90, 89, 60), Draw(True, 0, 0), Draw(False, 0, 2)]
    return base


def instanced_step_mode(width=160, height=120, instances=5) -> Scene:
    """A per-instance vertex buffer (VertexStepMode::Instance, vertex.rs:176-199): colours come from buffer 1."""
    tri = np.array([[-0.9, -0.8, 0.5, 1.0], [-0.5, 0.1, 0.5, 1.0], [-0.1, -0.8, 0.5, 1.0]], dtype=np.float32)
    u = splitmix64_u01(instances * 4, 71).reshape(instances, 4)
    colors = np.ones((instances, 4), dtype=np.float32)
    colors[:, :3] = u[:, :3]
    params = np.zeros(20, dtype=np.float32)
    params[:16] = np.eye(4, dtype=np.float32).reshape(-1)
    params[16:20] = [0.3, 0.15, 0.05, 0.0]
    return Scene(
        name="instanced_step_mode", width=width, height=height, shader="features",
        vertex_layouts=[VertexBufferLayout(16, "vertex", [VertexAttribute(0, "float32x4", 0)]),
                        VertexBufferLayout(16, "instance", [VertexAttribute(1, "float32x4", 0)])],
        vertex_buffers=[tri.view(np.uint8).reshape(-1), colors.view(np.uint8).reshape(-1)],
        bindings={(0, 0): ("buffer", params.view(np.uint8).reshape(-1).copy())},
        draws=[Draw(False, 0, 3, 0, 0, instances)],
    )


def huge_triangles(width=300, height=200) -> Scene:
    """Triangles far larger than the framebuffer (heavy clipping, the all-tiles list) and slivers."""
    v = np.array([
        [-50.0, -40.0, 0.2, 1.0, 1, 0, 0, 1], [60.0, -45.0, 0.9, 1.0, 0, 1, 0, 1], [3.0, 70.0, 0.4, 1.0, 0, 0, 1, 1],
        [-1.0, -1.0, 0.5, 1.0, 1, 1, 0, 1], [1.0, -1.0, 0.5, 1.0, 0, 1, 1, 1], [-1.0, 1.0, 0.5, 1.0, 1, 0, 1, 1],
        [-0.99, 0.2, 0.1, 1.0, 1, 1, 1, 1], [0.99, 0.21, 0.1, 1.0, 0, 0, 0, 1], [0.99, 0.2, 0.1, 1.0, 1, 0, 0, 1],
        [0.3, -3.0, 0.3, 0.5, 0, 1, 0, 1], [0.31, 3.0, 0.3, 2.0, 0, 0, 1, 1], [0.3, 3.0, 0.3, 1.0, 1, 1, 0, 1],
        [-2.0, -2.0, -0.5, 1.0, 1, 0, 0, 1], [2.0, -2.0, 0.5, 1.0, 0, 1, 0, 1], [0.0, 2.0, 1.5, 1.0, 0, 0, 1, 1],
    ], dtype=np.float32)
    return Scene(
        name="huge_triangles", width=width, height=height, shader="hello_mesh",
        vertex_layouts=[_POS_COLOR_LAYOUT], vertex_buffers=[v.view(np.uint8).reshape(-1)],
        bindings={(0, 0): ("buffer", identity_matrix_bytes())}, draws=[Draw(False, 0, 15)],
    )


def odd_sizes(width=333, height=77) -> Scene:
    """A framebuffer that is not a multiple of the 32x32 tile, with a viewport larger than it."""
    s = random_triangles(width, height, count=150, seed=81)
    s.name = f"odd_sizes_{width}x{height}"
    s.viewport = (-20.0, -10.0, 400.0, 100.0, 0.0, 1.0)
    return s


def fuzz(seed: int) -> Scene:
    """A random scene for the parity fuzz test: random framebuffer size, topology, index format / base vertex / first
    index, cull state, depth state, viewport, scissor and load ops over random (partly clipped, partly tiny and layered)
    geometry.  Everything is derived from `seed` with numpy's PCG64, so the oracle and the device see the same bytes."""
    rng = np.random.default_rng(1000 + seed)
    width, height = int(rng.integers(1, 330)), int(rng.integers(1, 230))
    topology = str(rng.choice(["triangle-list"] * 5 + ["triangle-strip", "line-list", "line-strip", "point-list"]))
    nverts = int(rng.integers(3, 1200))
    u = rng.random((nverts, 8), dtype=np.float32)
    v = np.ones((nverts, 8), dtype=np.float32)
    spread = np.float32(rng.choice([0.6, 1.0, 1.6, 3.0]))
    v[:, 0] = (u[:, 0] * 2 - 1) * spread
    v[:, 1] = (u[:, 1] * 2 - 1) * spread
    v[:, 2] = u[:, 2] * np.float32(1.3) - np.float32(0.15)
    if rng.random() < 0.5:
        v[:, 3] = np.float32(0.3) + u[:, 3] * np.float32(1.4)
    v[:, 4:7] = u[:, 4:7]
    small = rng.random() < 0.5
    if small and topology == "triangle-list":
        # tiny triangles in depth layers (the regime of C3, where the hierarchical depth test engages)
        t = v[: nverts // 3 * 3].reshape(-1, 3, 8)
        t[:, 1:, 0:2] = t[:, :1, 0:2] + (t[:, 1:, 0:2] - t[:, :1, 0:2]) * np.float32(0.02)
        layer = (np.arange(t.shape[0]) * 4 // max(t.shape[0], 1)).astype(np.float32).reshape(-1, 1)
        if rng.random() < 0.5:
            layer = layer[::-1]
        t[:, :, 2] = (layer + np.float32(0.5) + np.float32(0.4) * (t[:, :, 2] - np.float32(0.5))) / np.float32(4.0)
        t[:, :, 3] = 1.0
    indexed = bool(rng.random() < 0.6)
    index_data, strip_fmt, draws = None, None, []
    if indexed:
        fmt = np.uint16 if rng.random() < 0.5 else np.uint32
        nidx = int(rng.integers(3, 2500))
        base_vertex = int(rng.integers(0, max(nverts // 4, 1)))
        idx = rng.integers(0, nverts - base_vertex, nidx).astype(fmt)
        if small and topology == "triangle-list":      # keep the tiny triangles: consecutive vertices
            k = (nverts - base_vertex) // 3 * 3
            idx = np.resize(np.arange(k, dtype=fmt), nidx if k == 0 else max(nidx // 3 * 3, 3)) if k >= 3 else idx
        if topology.endswith("strip") and rng.random() < 0.6:
            strip_fmt = "uint16" if fmt == np.uint16 else "uint32"
            cut = rng.random(idx.size) < 0.08
            idx[cut] = np.iinfo(fmt).max
        first = int(rng.integers(0, max(idx.size // 5, 1)))
        index_data = idx
        draws.append(Draw(True, first, int(idx.size - first), base_vertex, 0, int(rng.integers(1, 3))))
    else:
        first = int(rng.integers(0, max(nverts // 5, 1)))
        draws.append(Draw(False, first, nverts - first, 0, 0, 1))
    compare = str(rng.choice(["less", "less", "less-equal", "greater", "greater-equal", "always", "never", "equal", "not-equal", "none"]))
    write = bool(rng.random() < 0.75)
    if compare == "not-equal" and not topology.startswith("triangle"):
        write = False       # keeps the seeds' scenes as they were when the ordered kernel rasterised triangles only
    s = Scene(
        name=f"fuzz_{seed}", width=width, height=height, shader="hello_mesh", topology=topology, strip_index_format=strip_fmt,
        front_face=str(rng.choice(["ccw", "cw"])), cull_mode=rng.choice([None, None, "front", "back"]),
        depth_compare=None if compare == "none" else compare, depth_write=write,
        clear_depth=float(rng.choice([1.0, 0.5, 0.25])),
        vertex_layouts=[_POS_COLOR_LAYOUT], vertex_buffers=[np.ascontiguousarray(v).view(np.uint8).reshape(-1)],
        index_data=index_data, bindings={(0, 0): ("buffer", identity_matrix_bytes())}, draws=draws,
    )
    if s.cull_mode is not None:
        s.cull_mode = str(s.cull_mode)
    if rng.random() < 0.3:
        s.viewport = (float(rng.integers(-20, 40)), float(rng.integers(-20, 40)), float(rng.integers(1, width + 60)),
                      float(rng.integers(1, height + 60)), 0.0, 1.0)
    if rng.random() < 0.3:
        x0, y0 = int(rng.integers(0, width)), int(rng.integers(0, height))
        s.scissor = (x0, y0, int(rng.integers(0, width - x0 + 1)), int(rng.integers(0, height - y0 + 1)))
    if rng.random() < 0.25:
        s.clear_color = None
        s.initial_color = rng.integers(0, 255, (height, width, 4), dtype=np.uint8)
        if s.depth_compare is not None:
            s.clear_depth = None
            s.initial_depth = rng.random((height, width), dtype=np.float32)
    return s


def fuzz_shaders(seed: int) -> Scene:
    """The state x geometry of fuzz(seed), drawn with one of the other parity shaders: instancing + flat varyings + discard,
    frag_depth, the early-depth-test programs (with depth writes, so that every topology reaches the ordered kernel),
    three colour attachments, perspective-correct varyings, primitive_index."""
    rng = np.random.default_rng(50000 + seed)
    s = fuzz(seed)
    shader = str(rng.choice(["features", "frag_depth", "early_force", "early_allow", "mrt", "perspective", "prim_index", "depth_only"]))
    s.name, s.shader = f"fuzz_{shader}_{seed}", shader
    s.color_format = "rgba8unorm"
    if s.initial_color is not None:                     # LoadOp::Load targets start from these bytes
        s.initial_color = np.ascontiguousarray(s.initial_color)
    if shader == "features":
        params = np.zeros(20, dtype=np.float32)
        params[:16] = np.eye(4, dtype=np.float32).reshape(-1)
        params[16:20] = [float(rng.random() * 0.3 - 0.15), float(rng.random() * 0.3 - 0.15), float(rng.random() * 0.1), 0.0]
        s.bindings = {(0, 0): ("buffer", params.view(np.uint8).reshape(-1).copy())}
        s.draws = [Draw(d.indexed, d.first, d.count, d.base_vertex, int(rng.integers(0, 3)), int(rng.integers(1, 4))) for d in s.draws]
    if shader in ("early_force", "early_allow", "depth_only") and s.depth_compare is not None:
        s.depth_write = True
    if shader == "mrt":
        s.extra_targets = [("bgra8unorm", (0.25, 0.5, 0.75, 1.0) if rng.random() < 0.7 else None), ("rg8unorm", (0.5, 0.125, 0.0, 0.0))]
    return s


_BLEND_FACTORS = ["zero", "one", "src", "one-minus-src", "src-alpha", "one-minus-src-alpha", "dst", "one-minus-dst", "dst-alpha",
                  "one-minus-dst-alpha", "src-alpha-saturated", "constant", "one-minus-constant"]


def fuzz_features(seed: int) -> Scene:
    """fuzz(seed) on a device with a random set of the opt-in features (viewport depth range, colour write mask, blending
    with a random blend state) over a random non-sRGB colour format.  The sRGB encode is left out: it goes through powf on
    both sides and is held to a tolerance, not to equality (tests/test_features_gpu.py)."""
    rng = np.random.default_rng(70000 + seed)
    s = fuzz(seed)
    s.name = f"fuzz_features_{seed}"
    s.color_format = str(rng.choice(["rgba8unorm", "bgra8unorm"]))
    v = s.vertex_buffers[0].view(np.float32).reshape(-1, 8).copy()
    v[:, 7] = rng.random(v.shape[0], dtype=np.float32)                  # translucent vertices
    s.vertex_buffers[0] = v.view(np.uint8).reshape(-1)
    s.features = 0
    if rng.random() < 0.6:
        s.features |= 16                                                # WGB_FEATURE_BLEND
        ops = ["add", "subtract", "reverse-subtract", "min", "max"]
        pick = lambda: (str(rng.choice(_BLEND_FACTORS)), str(rng.choice(_BLEND_FACTORS)), str(rng.choice(ops)))
        s.blend = {"color": pick(), "alpha": pick()}
        s.blend_constant = tuple(float(x) for x in rng.random(4))
    if rng.random() < 0.5:
        s.features |= 2                                                 # WGB_FEATURE_COLOR_WRITE_MASK
        s.color_write_mask = int(rng.integers(0, 16))
    if rng.random() < 0.4:
        s.features |= 1                                                 # WGB_FEATURE_VIEWPORT_DEPTH_RANGE
        lo = float(rng.random() * 0.5)
        vp = s.viewport or (0.0, 0.0, float(s.width), float(s.height), 0.0, 1.0)
        s.viewport = (vp[0], vp[1], vp[2], vp[3], lo, lo + float(rng.random() * 0.5))
    if s.clear_color is not None:
        s.clear_color = tuple(float(x) for x in rng.random(4))
    return s


def fuzz_textured(seed: int) -> Scene:
    """Random textured triangles for the sampling path (binding.rs:93-164): texture coordinates from far outside [0, 1] to
    exactly on texel boundaries, every address mode per axis, texture sizes from 1 x 1 over odd and non-square ones,
    both sampled formats, perspective division (w != 1) in the positions."""
    rng = np.random.default_rng(90000 + seed)
    width, height = int(rng.integers(8, 200)), int(rng.integers(8, 150))
    count = int(rng.integers(1, 120))
    u = rng.random((count * 3, 6), dtype=np.float32)
    v = np.ones((count * 3, 6), dtype=np.float32)
    spread = np.float32(rng.choice([0.7, 1.0, 1.5]))
    v[:, 0] = (u[:, 0] * 2 - 1) * spread
    v[:, 1] = (u[:, 1] * 2 - 1) * spread
    v[:, 2] = u[:, 2]
    if rng.random() < 0.5:
        v[:, 3] = np.float32(0.4) + u[:, 3] * np.float32(1.2)
    scale = np.float32(rng.choice([1.0, 1.0, 3.0, 17.0]))
    v[:, 4:6] = (u[:, 4:6] * 2 - np.float32(rng.choice([0.0, 1.0]))) * scale
    tw, th = int(rng.choice([1, 2, 3, 7, 16, 33, 64])), int(rng.choice([1, 2, 5, 8, 31, 64]))
    if rng.random() < 0.3:       # coordinates that land exactly on texel centres and edges of the texture
        v[:, 4] = np.round(v[:, 4] * np.float32(2 * max(tw - 1, 1))) / np.float32(2 * max(tw - 1, 1))
        v[:, 5] = np.round(v[:, 5] * np.float32(2 * max(th - 1, 1))) / np.float32(2 * max(th - 1, 1))
    image = rng.integers(0, 256, (th, tw, 4), dtype=np.uint8)
    modes = ["clamp-to-edge", "repeat", "mirror-repeat"]
    return Scene(
        name=f"fuzz_textured_{seed}", width=width, height=height, shader="hello_texture", color_format=str(rng.choice(["rgba8unorm", "rgba8unorm-srgb"])),
        front_face="ccw", cull_mode=None, depth_compare=str(rng.choice(["less", "always", "greater-equal"])), depth_write=True,
        clear_depth=float(rng.choice([1.0, 0.5])),
        vertex_layouts=[_POS_UV_LAYOUT], vertex_buffers=[np.ascontiguousarray(v).view(np.uint8).reshape(-1)],
        bindings={(0, 0): ("buffer", identity_matrix_bytes()), (1, 0): ("texture", image, str(rng.choice(["rgba8unorm", "rgba8unorm-srgb"]))),
                  (1, 1): ("sampler", str(rng.choice(modes)), str(rng.choice(modes)))},
        draws=[Draw(False, 0, count * 3)],
    )
