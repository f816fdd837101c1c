"""GPU tests written after round 1's GPU minutes were spent: they pass on the software model of tests/cusim (the CPU
tier runs them there) and their kernels compile for sm_100a, but they have not run on a B200 yet.  The file sorts last so
that on the hardware they run after every test that has.  (Same helpers and bar as tests/test_parity_gpu.py: coverage,
depth bits and colour bytes equal to the oracle's.)"""
import numpy as np
import pytest

from tests.test_features_gpu import BLENDS, _render_both, _translucent, ext  # noqa: F401  (ext is a fixture)
from tests.test_parity_gpu import _compare, gpu  # noqa: F401  (gpu is a fixture)
from wgpu_cpu_b200 import scenes as S

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,compare,topology", [
    ("force", "less", "triangle-list"), ("force", "greater-equal", "triangle-list"), ("force", "not-equal", "triangle-list"),
    ("allow", "less", "triangle-list"), ("allow", "less-equal", "triangle-list"), ("allow", "greater-equal", "triangle-list"),
    ("force", "less", "line-list"), ("allow", "less-equal", "line-strip"), ("force", "less-equal", "point-list")])
def test_early_depth_test_before_discard_and_frag_depth(gpu, kind, compare, topology):
    """fragment.rs:166-194: with @early_depth_test the rasteriser's depth is tested and written before the stage runs --
    a fragment the stage then discards has already left its depth behind, the frag_depth it returns is ignored (force)
    or tested once more against the depth just stored (allow).  The fold has no closed form:
    these pipelines take the ordered tile kernel."""
    got, ref = _compare(S.early_depth(kind, compare, topology=topology), gpu)
    if compare != "not-equal":
        assert ref.stats["fragments_shaded"] < got.stats["fragments"]          # the early test rejected fragments before the stage ran


@pytest.mark.parametrize("topology", ["line-list", "line-strip", "point-list"])
def test_not_equal_with_depth_write_on_lines_and_points(gpu, topology):
    """The ordered kernel walks lines once per thread and treats points as single pixels."""
    s = S.random_lines(200, 150, 90, 12, topology)
    s.depth_compare, s.depth_write, s.clear_depth = "not-equal", True, 0.5
    _compare(s, gpu)


def test_wgb_log_prints_the_pass_timer_and_counters():
    """The reference's only instrumentation is `tracing::debug!(?elapsed, "render pass time")` and its per-draw counters
    (render_pass/mod.rs:346,392-393; state.rs:516-517,592).  WGB_LOG=debug prints the same per pass, WGB_LOG=trace per
    draw as well (read once per process, hence the child process)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import os, sys; sys.path.insert(0, %r)\n"
            "if os.environ.get('WGB_CUSIM') == '1':\n    from tests import conftest; conftest.use_cusim()\n"
            "from wgpu_cpu_b200 import api, scenes\nfrom wgpu_cpu_b200.render import render_scene\n"
            "dev, q = api.instance().request_adapter().request_device(0)\nrender_scene(dev, q, scenes.multi_draw())\n") % root
    p = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, WGB_LOG="trace"), capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    lines = [ln for ln in p.stderr.splitlines() if ln.startswith("wgpu-b200 ")]
    assert any("DEBUG render pass time:" in ln and "primitives_drawn=89" in ln and "draws=3" in ln for ln in lines), p.stderr
    assert sum("TRACE draw" in ln for ln in lines) >= 3
    quiet = subprocess.run([sys.executable, "-c", code], env={k: v for k, v in os.environ.items() if k != "WGB_LOG"}, capture_output=True, text=True, timeout=600)
    assert quiet.returncode == 0 and "wgpu-b200 " not in quiet.stderr


def test_concurrent_recording_and_submission_from_threads(gpu):
    """Backend objects are Send + Sync and recording may happen on any thread (engine.rs:26-36 executes the
    submissions one at a time in submission order).  Four threads share one device and queue; each creates its own
    resources and pipeline, records, submits and reads back three times (ctypes drops the GIL around every C call)."""
    import threading
    from oracle import pyoracle
    from wgpu_cpu_b200.render import SceneRenderer
    dev, queue = gpu
    scn = [S.random_triangles(count=300, seed=40), S.hello_mesh(200, 150), S.random_lines(160, 120, 80, 7, "line-strip"), S.features()]
    refs = [pyoracle.render(s, want_coverage=False) for s in scn]
    errors = []

    def work(i):
        try:
            for _ in range(3):
                r = SceneRenderer(dev, queue, scn[i])
                r.render()
                f = r.read()
                if not np.array_equal(f.color, refs[i].color) or (refs[i].depth is not None and not np.array_equal(f.depth.view(np.uint32), refs[i].depth.view(np.uint32))):
                    errors.append(f"{scn[i].name}: frame differs")
        except Exception as e:      # noqa: BLE001 -- reported below, on the main thread
            errors.append(f"{scn[i].name}: {e!r}")

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(scn))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


@pytest.mark.parametrize("compare", ["less", "not-equal"])
def test_clip_record_and_big_list_overflow_replay(monkeypatch, compare):
    """Clip records and the big list live in fixed buffers; when either overflows the tile kernel leaves the attachments
    untouched and the host replays the draw with larger ones.  WGB_TEST_SMALL_WORK_BUFFERS=1 starts both at two entries, so
    a scene with clipped and tile-spanning triangles overflows both, repeatedly (closed-form and ordered tile kernel)."""
    from oracle import pyoracle
    from wgpu_cpu_b200 import api
    from wgpu_cpu_b200.render import SceneRenderer
    monkeypatch.setenv("WGB_TEST_SMALL_WORK_BUFFERS", "1")
    dev, queue = api.instance().request_adapter().request_device(0)
    monkeypatch.delenv("WGB_TEST_SMALL_WORK_BUFFERS")
    scene = S.huge_triangles()
    scene.depth_compare, scene.depth_write = compare, True
    ref = pyoracle.render(scene, want_coverage=False)
    r = SceneRenderer(dev, queue, scene)
    st = r.render()
    assert st["replays"] >= 2 and st["big_primitives"] > 2 and st["clip_records"] > 2
    f = r.read()
    assert np.array_equal(f.color, ref.color) and np.array_equal(f.depth.view(np.uint32), ref.depth.view(np.uint32))
    assert r.render()["replays"] == 0                       # the device keeps the capacities it learned
    assert np.array_equal(r.read().color, ref.color)


@pytest.mark.parametrize("topology", ["line-list", "line-strip", "point-list"])
def test_blending_lines_and_points(ext, topology):
    """Blended lines and points: the ordered kernel runs every topology (one Bresenham walk per thread for a line)."""
    from wgpu_cpu_b200 import api
    scene = S.random_lines(200, 150, 120, 17, topology)
    scene.color_format = "rgba8unorm"
    _translucent(scene, 4)
    scene.features = api.FEATURE["BLEND"]
    scene.blend = BLENDS["alpha"]
    scene.clear_color = (0.1, 0.3, 0.2, 0.5)
    got, ref = _render_both(scene, ext)
    assert np.array_equal(got.color, ref.color)
    scene.features = 0
    from oracle import pyoracle
    assert not np.array_equal(pyoracle.render(scene, want_coverage=False).color, ref.color)


@pytest.mark.parametrize("config", ["c1", "c2", "c3", "c4", "c5_frame_17"])
def test_baseline_configs_at_full_size(gpu, config):
    """BASELINE.json's configurations at their full sizes, bit-exact against the oracle (colour bytes and depth bits of
    every pixel): C1 teapot 512x512, C2 bunny 1920x1080 textured, C3 9 999 392 triangles at 3840x2160, C4 the 64-iteration
    fragment shader at 7680x4320, and one frame (yaw 17/64 of a turn) of C5's bunny at 3840x2160.  The oracle needs a few
    seconds for each; the production configuration is what is measured (no coverage capture, hierarchical depth test on)."""
    import math
    from oracle import pyoracle
    from wgpu_cpu_b200.render import render_scene
    dev, queue = gpu
    scene = {"c1": lambda: S.hello_mesh(512, 512), "c2": lambda: S.hello_texture(1920, 1080), "c3": lambda: S.synthetic_grid(),
             "c4": lambda: S.procedural(7680, 4320),
             "c5_frame_17": lambda: S.hello_texture(3840, 2160, yaw=17 * 2.0 * math.pi / 64)}[config]()
    ref = pyoracle.render(scene, want_coverage=False)
    got = render_scene(dev, queue, scene, want_coverage=False)
    assert np.array_equal(got.color, ref.color), f"{scene.name}: colour differs at {int((got.color != ref.color).any(axis=2).sum())} pixels"
    if ref.depth is not None:
        assert np.array_equal(got.depth.view(np.uint32), ref.depth.view(np.uint32)), f"{scene.name}: depth differs"
    assert got.stats["primitives"] == scene.num_primitives
    if config == "c3":
        assert got.stats["primitives"] == 9999392 and got.stats["bin_pairs"] == 11952651      # the figures of profiles/r01_summary.md


def test_integer_vector_vertex_formats(gpu):
    """Uint32x2..x4 / Sint32x2..x4 attributes: the vertex stage receives the attribute's raw bytes as its declared type
    (vertex.rs:128-156 copies `format.size()` bytes), integers travel to the fragment stage as flat varyings."""
    from wgpu_cpu_b200 import api
    dev, queue = gpu
    wgsl = """
struct VOut { @builtin(position) p: vec4f, @location(0) @interpolate(flat) packed: vec4u, @location(1) @interpolate(flat) s: vec2i, }
struct FOut { @builtin(frag_depth) depth: f32, @location(0) color: vec4f, }
@vertex fn vs_main(@builtin(vertex_index) i: u32, @location(0) a: vec4u, @location(1) b: vec2i, @location(2) c: vec3u) -> VOut {
    let x = f32(i32(i & 1u) * 4 - 1);
    let y = f32(i32(i >> 1u) * 4 - 1);
    return VOut(vec4f(x, y, 0.5, 1.0), vec4u(a.x + c.x, a.y + c.y, a.z + c.z, a.w), b);
}
@fragment fn fs_main(v: VOut) -> FOut {
    let sum = f32(v.packed.x) + f32(v.packed.y) * 10.0 + f32(v.packed.z) * 100.0 + f32(v.packed.w) * 1000.0 + f32(v.s.x * v.s.y);
    return FOut(sum / 65536.0, vec4f(1.0, 0.0, 0.0, 1.0));
}"""
    module = dev.create_shader_module(wgsl)
    # one 48-byte vertex, three copies: a = (1,2,3,4) @0, b = (-3, 5) @16, c = (4,3,2) @24
    vertex = np.zeros(12, dtype=np.uint32)
    vertex[0:4] = [1, 2, 3, 4]
    vertex[4:6] = np.array([-3, 5], dtype=np.int32).view(np.uint32)
    vertex[6:9] = [4, 3, 2]
    vb = dev.create_buffer_init(np.tile(vertex, 3), api.BUFFER_USAGE["VERTEX"])
    pipe = dev.create_render_pipeline(
        vertex_module=module, fragment_module=module, targets=["rgba8unorm"],
        vertex_buffers=[{"array_stride": 48, "attributes": [("uint32x4", 0, 0), ("sint32x2", 16, 1), ("uint32x3", 24, 2)]}],
        depth_stencil={"depth_compare": "always", "depth_write_enabled": True})
    color = dev.create_texture(8, 8, "rgba8unorm")
    depth = dev.create_texture(8, 8, "depth32float")
    enc = dev.create_command_encoder()
    with enc.begin_render_pass([{"view": color.create_view(), "load": ("clear", (0, 0, 0, 0))}],
                               {"view": depth.create_view(), "depth_load": ("clear", 0.0)}) as rp:
        rp.set_pipeline(pipe)
        rp.set_vertex_buffer(0, vb)
        rp.draw(range(0, 3))
    dev.poll(True, queue.submit([enc.finish()]))
    d = depth.read()
    assert (color.read()[..., 0] == 255).all()
    want = np.float32((5 + 5 * 10 + 5 * 100 + 4 * 1000 - 15) / 65536.0)
    assert (d == want).all(), (d[0, 0], want)


def test_explicit_layouts_and_stream_handle(gpu):
    """hello_mesh.rs:126-142,174-180 creates a bind group layout and a pipeline layout and hands them to the pipeline and
    the bind group; the frame is the same as with the layouts left out.  The device's render stream is exposed for
    interop (NCCL presenter, CUDA-event timing)."""
    from oracle import pyoracle
    from wgpu_cpu_b200 import api, shaders
    dev, queue = gpu
    scene = S.hello_mesh(96, 64)
    ref = pyoracle.render(scene, want_coverage=False)
    module = dev.create_shader_module(shaders.wgsl("hello_mesh"))
    bgl = dev.create_bind_group_layout([(0, 1, 1, 0)])                       # binding 0, vertex stage, buffer, no dynamic offset
    layout = dev.create_pipeline_layout([bgl])
    pipe = dev.create_render_pipeline(
        vertex_module=module, fragment_module=module, layout=layout, front_face=scene.front_face, cull_mode=scene.cull_mode,
        vertex_buffers=[{"array_stride": 32, "attributes": [("float32x4", 0, 0), ("float32x4", 16, 1)]}],
        depth_stencil={"depth_compare": "less", "depth_write_enabled": True}, targets=[scene.color_format])
    vb = dev.create_buffer_init(scene.vertex_buffers[0], api.BUFFER_USAGE["VERTEX"])
    ib = dev.create_buffer_init(scene.index_data, api.BUFFER_USAGE["INDEX"])
    ub = dev.create_buffer_init(scene.bindings[(0, 0)][1], api.BUFFER_USAGE["UNIFORM"])
    group = dev.create_bind_group(bgl, [{"binding": 0, "buffer": ub}])
    color = dev.create_texture(scene.width, scene.height, scene.color_format)
    depth = dev.create_texture(scene.width, scene.height, "depth32float")
    enc = dev.create_command_encoder()
    with enc.begin_render_pass([{"view": color.create_view(), "load": ("clear", scene.clear_color)}],
                               {"view": depth.create_view(), "depth_load": ("clear", scene.clear_depth)}) as rp:
        rp.set_pipeline(pipe)
        rp.set_bind_group(0, group)
        rp.set_vertex_buffer(0, vb)
        rp.set_index_buffer(ib, "uint32")
        d = scene.draws[0]
        rp.draw_indexed(range(d.first, d.first + d.count), d.base_vertex, range(d.first_instance, d.first_instance + d.instance_count))
    dev.poll(True, queue.submit([enc.finish()]))
    assert np.array_equal(color.read(), ref.color)
    assert np.array_equal(depth.read().view(np.uint32), ref.depth.view(np.uint32))
    assert dev.stream() != 0


@pytest.mark.parametrize("compare,second_load", [("less", False), ("greater-equal", False), ("less", True), ("not-equal", False)])
def test_multiple_colour_targets(gpu, compare, second_load):
    """Three colour attachments (rgba8unorm, bgra8unorm, rg8unorm) written by one fragment stage whose outputs are
    declared out of location order (fragment.rs:457-488): every attachment bit-exact, with a cleared and a loaded second
    attachment, through the closed-form tile kernel and (NotEqual + write) the ordered one."""
    from oracle import pyoracle
    from wgpu_cpu_b200.render import SceneRenderer
    dev, queue = gpu
    scene = S.multiple_targets(compare=compare, second_load=second_load)
    ref = pyoracle.render(scene, want_coverage=False)
    r = SceneRenderer(dev, queue, scene)
    r.render()
    got = r.read()
    assert np.array_equal(got.color, ref.color)
    assert np.array_equal(got.depth.view(np.uint32), ref.depth.view(np.uint32))
    assert len(got.extra_colors) == 2
    for k, (g, e) in enumerate(zip(got.extra_colors, ref.extra_colors), start=1):
        assert g.shape == e.shape and np.array_equal(g, e), f"colour attachment {k} differs at {int((g != e).any(axis=2).sum())} pixels"
    assert ref.extra_colors[0].any() and ref.extra_colors[1][..., 1].any()


def test_negative_base_vertex_and_first_instance(gpu):
    """index.rs:45-62: vertex = index + base_vertex with a signed base (indices shifted up by 100, base_vertex = -100), and
    a draw whose instance range starts at 2 (instance_index reaches the vertex stage as first_instance + k)."""
    s = S.features()
    base = s.vertex_buffers[0]
    s.index_data = (np.arange(180, dtype=np.uint32) + 100)
    s.draws = [S.Draw(True, 0, 180, -100, 2, 2)]
    _compare(s, gpu)
    s.draws = [S.Draw(True, 0, 180, -101, 0, 1)]          # index 100 - 101 underflows: strict_add_signed panics in the reference
    from wgpu_cpu_b200 import api
    from wgpu_cpu_b200.render import render_scene
    with pytest.raises(api.WgpuError):
        render_scene(gpu[0], gpu[1], s)
    assert base is s.vertex_buffers[0]


def test_bindings_in_higher_groups_and_array_layers(gpu):
    """Resources in bind groups 1 and 2 (binding.rs:22-52 flattens (group, binding)), and a render pass into array layer 1
    of a two-layer texture through a view with base_array_layer = 1: layer 0 keeps its bytes."""
    from wgpu_cpu_b200 import api
    dev, queue = gpu
    wgsl = """
struct A { scale: vec4f, }
struct B { bias: vec4f, }
@group(1) @binding(3) var<uniform> a: A;
@group(2) @binding(0) var<uniform> b: B;
@vertex fn vs_main(@builtin(vertex_index) i: u32) -> @builtin(position) vec4f {
    return vec4f(f32(i32(i & 1u) * 4 - 1), f32(i32(i >> 1u) * 4 - 1), 0.5, 1.0);
}
@fragment fn fs_main() -> @location(0) vec4f {
    return vec4f(a.scale.x * b.bias.x, a.scale.y + b.bias.y, a.scale.z - b.bias.z, 1.0);
}"""
    m = dev.create_shader_module(wgsl)
    pipe = dev.create_render_pipeline(vertex_module=m, fragment_module=m, targets=["rgba8unorm"])
    ua = dev.create_buffer_init(np.array([0.5, 0.25, 0.75, 0.0], dtype=np.float32), api.BUFFER_USAGE["UNIFORM"])
    ub = dev.create_buffer_init(np.array([0.5, 0.5, 0.25, 0.0], dtype=np.float32), api.BUFFER_USAGE["UNIFORM"])
    g1 = dev.create_bind_group(None, [{"binding": 3, "buffer": ua}])
    g2 = dev.create_bind_group(None, [{"binding": 0, "buffer": ub}])
    tex = dev.create_texture(40, 24, "rgba8unorm", layers=2)
    enc = dev.create_command_encoder()
    with enc.begin_render_pass([{"view": tex.create_view(base_array_layer=1), "load": ("clear", (0, 0, 0, 0))}], None) as rp:
        rp.set_pipeline(pipe)
        rp.set_bind_group(1, g1)
        rp.set_bind_group(2, g2)
        rp.draw(range(0, 3))
    dev.poll(True, queue.submit([enc.finish()]))
    texels = tex.read()
    assert texels.shape == (2, 24, 40, 4)
    assert not texels[0].any()
    want = [int(np.float32(0.25) * np.float32(255.0)), int(np.float32(0.75) * np.float32(255.0)), int(np.float32(0.5) * np.float32(255.0)), 255]
    assert (texels[1] == np.array(want, dtype=np.uint8)).all(), texels[1][0, 0]


def test_several_command_buffers_and_submissions_without_waiting(gpu):
    """Submissions execute in submission order (engine.rs:26-36) whether or not the application waits in between: three
    scenes submitted back to back (two command buffers in one submit, one in the next), one poll(Wait) at the end."""
    from oracle import pyoracle
    from wgpu_cpu_b200.render import SceneRenderer
    dev, queue = gpu
    scn = [S.random_triangles(count=150, seed=51), S.random_lines(160, 120, 50, 52, "line-list"), S.hello_mesh(120, 90)]
    rs = [SceneRenderer(dev, queue, s) for s in scn]
    first = queue.submit([rs[0].encode(), rs[1].encode()])
    second = queue.submit([rs[2].encode()])
    assert second == first + 1                                  # device.rs:443-444: monotonically increasing
    dev.poll(True, second)
    for r, s in zip(rs, scn):
        ref = pyoracle.render(s, want_coverage=False)
        f = r.read()
        assert np.array_equal(f.color, ref.color), s.name
        if ref.depth is not None:
            assert np.array_equal(f.depth.view(np.uint32), ref.depth.view(np.uint32)), s.name


@pytest.mark.parametrize("compare", ["less", "greater"])
def test_depth_only_pass(gpu, compare):
    """A fragment stage without outputs never reaches the late depth test (it runs at the first @location output,
    fragment.rs:457-488); with @early_depth_test(force) the early test (fragment.rs:166-194) tests and writes depth.
    Rendered once next to an (untouched) colour attachment and once as a true depth-only pass."""
    from oracle import pyoracle
    from wgpu_cpu_b200 import api, shaders
    dev, queue = gpu
    scene = S.random_triangles(200, 140, count=150, seed=71, spread=1.1, with_w=False)
    scene.name, scene.shader = f"depth_only_{compare}", "depth_only"
    scene.depth_compare, scene.depth_write, scene.clear_depth = compare, True, (1.0 if compare == "less" else 0.25)
    got, ref = _compare(scene, gpu)
    assert (ref.color == ref.color[0, 0]).all() and (ref.depth != scene.clear_depth).any()
    # no colour attachment at all
    module = dev.create_shader_module(shaders.wgsl("depth_only"))
    pipe = dev.create_render_pipeline(
        vertex_module=module, fragment_module=module, targets=[], front_face=scene.front_face, cull_mode=scene.cull_mode,
        vertex_buffers=[{"array_stride": 32, "attributes": [("float32x4", 0, 0), ("float32x4", 16, 1)]}],
        depth_stencil={"depth_compare": compare, "depth_write_enabled": True})
    vb = dev.create_buffer_init(scene.vertex_buffers[0], api.BUFFER_USAGE["VERTEX"])
    ub = dev.create_buffer_init(scene.bindings[(0, 0)][1], api.BUFFER_USAGE["UNIFORM"])
    depth = dev.create_texture(scene.width, scene.height, "depth32float")
    enc = dev.create_command_encoder()
    with enc.begin_render_pass([], {"view": depth.create_view(), "depth_load": ("clear", scene.clear_depth)}) as rp:
        rp.set_pipeline(pipe)
        rp.set_bind_group(0, dev.create_bind_group(None, [{"binding": 0, "buffer": ub}]))
        rp.set_vertex_buffer(0, vb)
        d = scene.draws[0]
        rp.draw(range(d.first, d.first + d.count))
    dev.poll(True, queue.submit([enc.finish()]))
    assert np.array_equal(depth.read().view(np.uint32), ref.depth.view(np.uint32))


@pytest.mark.parametrize("kind", ["list_culled_clipped_instanced", "strip_with_restart", "lines", "not_equal_ordered"])
def test_primitive_index_sample_index_and_mask(gpu, kind):
    """@builtin(primitive_index) is the primitive's position in its instance's assembly order, counted before culling and
    clipping and restarting with every instance (state.rs:519-541); sample_index is 0 and sample_mask all ones."""
    if kind == "list_culled_clipped_instanced" or kind == "not_equal_ordered":
        s = S.random_triangles(220, 160, count=400, seed=81, spread=1.4)
        s.cull_mode, s.front_face = "back", "ccw"
        s.draws = [S.Draw(False, 0, 1200, 0, 0, 2)]
        if kind == "not_equal_ordered":
            s.depth_compare, s.depth_write, s.clear_depth = "not-equal", True, 0.5
    elif kind == "lines":
        s = S.random_lines(200, 150, 300, 82, "line-strip")
    elif kind == "strip_with_restart":
        rng = np.random.default_rng(83)
        s = S.random_triangles(220, 160, count=200, seed=83, spread=1.2, with_w=False)
        idx = rng.integers(0, 600, 900).astype(np.uint16)
        idx[rng.random(900) < 0.1] = 0xFFFF
        s.topology, s.strip_index_format, s.index_data = "triangle-strip", "uint16", idx
        s.draws = [S.Draw(True, 0, 900, 0, 0, 1)]
    s.name, s.shader = f"prim_index_{kind}", "prim_index"
    got, ref = _compare(s, gpu)
    assert len(np.unique(ref.color[..., 0])) > 20 and (ref.color[..., 2] == 127).any()


@pytest.mark.parametrize("topology", ["triangle-list", "line-list", "point-list"])
def test_perspective_correct_interpolation(gpu, topology):
    """WGSL's default interpolation is perspective-correct; the reference panics on it (fragment.rs:343-345 todo!()), so the
    oracle restates WebGPU's formula -- sum(v_k B_k / w_k) / sum(B_k / w_k) -- and the frame must match it bit for bit,
    clipped primitives included.  A linear and a flat varying ride along; the corrected one must differ from the linear."""
    if topology == "triangle-list":
        s = S.random_triangles(220, 160, count=160, seed=91, spread=1.3, with_w=True)
    else:
        s = S.random_lines(220, 160, 120, 92, topology)
    s.name, s.shader = f"perspective_{topology}", "perspective"
    got, ref = _compare(s, gpu)
    if topology != "point-list":
        # red = perspective-corrected tint.x, blue = the same value interpolated linearly in screen space
        assert (ref.color[..., 0].astype(int) != ref.color[..., 2].astype(int)).sum() > 200


@pytest.mark.parametrize("name,blend", [
    ("src_factors", {"color": ("src", "one-minus-src", "add"), "alpha": ("one-minus-dst-alpha", "one-minus-dst", "add")}),
    ("dst_factors", {"color": ("one-minus-dst", "src", "subtract"), "alpha": ("one-minus-src", "one-minus-dst-alpha", "reverse-subtract")})])
def test_blend_factors_not_covered_elsewhere(ext, name, blend):
    """Src, OneMinusSrc, OneMinusDst and OneMinusDstAlpha (the kernel-coverage run of tests/cusim showed that no other
    blending case selects them)."""
    from wgpu_cpu_b200 import api
    scene = S.random_triangles(count=200, seed=6, color_format="bgra8unorm", depth_compare=None, depth_write=False)
    _translucent(scene, 10)
    scene.features = api.FEATURE["BLEND"]
    scene.blend = blend
    scene.clear_color = (0.2, 0.6, 0.4, 0.7)
    got, ref = _render_both(scene, ext)
    assert np.array_equal(got.color, ref.color)


def test_clear_only_pass_and_ordered_kernel_with_exact_bins(gpu, monkeypatch):
    """A render pass without a draw still applies its LoadOp::Clear (State::load, state.rs:135-145): that is the clear
    kernel, which nothing else reaches.  And the ordered tile kernel over exact-size bins (count / scan / fill instead of
    direct binning, WGB_NO_DIRECT_BINS=1)."""
    from oracle import pyoracle
    from wgpu_cpu_b200 import api
    from wgpu_cpu_b200.render import render_scene
    dev, queue = gpu
    for w, h in ((70, 45), (64, 64)):
        color = dev.create_texture(w, h, "bgra8unorm")
        extra = dev.create_texture(w, h, "rg8unorm")
        depth = dev.create_texture(w, h, "depth32float")
        enc = dev.create_command_encoder()
        with enc.begin_render_pass([{"view": color.create_view(), "load": ("clear", (0.25, 0.5, 1.0, 0.0))},
                                    {"view": extra.create_view(), "load": ("clear", (1.0, 0.5, 0.0, 0.0))}],
                                   {"view": depth.create_view(), "depth_load": ("clear", 0.375)}):
            pass
        dev.poll(True, queue.submit([enc.finish()]))
        assert (color.read() == np.array([255, 127, 63, 0], dtype=np.uint8)).all()          # B G R A
        assert (extra.read() == np.array([255, 127], dtype=np.uint8)).all()
        assert (depth.read() == np.float32(0.375)).all()
    monkeypatch.setenv("WGB_NO_DIRECT_BINS", "1")
    dev2, queue2 = api.instance().request_adapter().request_device(0)
    monkeypatch.delenv("WGB_NO_DIRECT_BINS")
    scene = S.random_triangles(count=250, seed=77, clear_depth=0.5)
    scene.depth_compare, scene.depth_write = "not-equal", True
    ref = pyoracle.render(scene, want_coverage=False)
    got = render_scene(dev2, queue2, scene, want_coverage=False)
    assert np.array_equal(got.color, ref.color) and np.array_equal(got.depth.view(np.uint32), ref.depth.view(np.uint32))


def test_vertex_attributes_that_are_only_four_byte_aligned(gpu):
    """A 28-byte vertex (position, one pad word, uv) leaves the float32x2 attribute on 4-byte boundaries and the float32x4
    one on 4- or 8-byte ones for most vertices: the scalar fetch paths behind the 128-bit / 64-bit ones."""
    s = S.hello_texture(200, 150)
    v = s.vertex_buffers[0].view(np.float32).reshape(-1, 6)
    padded = np.zeros((v.shape[0], 7), dtype=np.float32)
    padded[:, 0:4], padded[:, 5:7] = v[:, 0:4], v[:, 4:6]
    s.vertex_buffers = [padded.view(np.uint8).reshape(-1)]
    s.vertex_layouts = [S.VertexBufferLayout(28, "vertex", [S.VertexAttribute(0, "float32x4", 0), S.VertexAttribute(1, "float32x2", 20)])]
    s.name = "hello_texture_unaligned_attributes"
    _compare(s, gpu)


@pytest.mark.parametrize("name", ["hello_mesh", "hello_texture", "features", "frag_depth"])
def test_pre_emitted_shader_modules(gpu, name):
    """wgb_shader_module_descriptor.emitted: a host that runs its own WGSL / naga-IR front end hands the CUDA C++ of each
    entry point to the library instead of WGSL.  The golden emitter output of the built-in shaders,
    fed back this way, renders the same frame."""
    scene = {"hello_mesh": lambda: S.hello_mesh(128, 96), "hello_texture": lambda: S.hello_texture(160, 100),
             "features": S.features, "frag_depth": S.frag_depth}[name]()
    _compare(scene, gpu, use_emitted=True)


def test_queue_write_buffer_ranges_timer_and_device_pointer(gpu, tmp_path):
    """Queue::write_buffer with an offset updates that range only (the reference's copy_from_slice on the whole buffer
    only works for whole-buffer writes, device.rs:341-343 / buffer.rs:297-309 -- the intended semantics are implemented);
    a frame rendered after the update sees it.  Also: the CUDA-event timer around submissions, the texel storage's
    device pointer, and dump_texture of a BGRA target."""
    from oracle import pyoracle
    from tests.test_c_abi import _decode_png
    from wgpu_cpu_b200.render import SceneRenderer
    dev, queue = gpu
    scene = S.random_triangles(160, 120, count=120, seed=95, color_format="bgra8unorm")
    r = SceneRenderer(dev, queue, scene)
    dev.timer_begin()
    r.render()
    assert dev.timer_end() > 0.0
    before = r.read().color
    assert np.array_equal(before, pyoracle.render(scene, want_coverage=False).color)
    # move the second half of the vertices: 60 triangles x 3 vertices x 32 bytes, written at their byte offset
    v = scene.vertex_buffers[0].view(np.float32).reshape(-1, 8).copy()
    v[180:, 0] = -v[180:, 0]
    queue.write_buffer(r.vertex_buffers[0], 180 * 32, v[180:])
    scene.vertex_buffers[0] = v.view(np.uint8).reshape(-1)
    r.render()
    after = r.read().color
    assert np.array_equal(after, pyoracle.render(scene, want_coverage=False).color) and not np.array_equal(after, before)
    ptr, nbytes = r.target.device_pointer()
    assert ptr != 0 and nbytes == 160 * 120 * 4
    r.target.dump_png(str(tmp_path / "bgra.png"))
    assert np.array_equal(_decode_png(str(tmp_path / "bgra.png")), after[..., [2, 1, 0, 3]])      # PNGs are RGBA
    with pytest.raises(Exception):
        queue.write_buffer(r.vertex_buffers[0], scene.vertex_buffers[0].nbytes - 16, np.zeros(32, dtype=np.uint8))   # past the end


@pytest.mark.parametrize("seed", range(16))
def test_fuzz_with_the_other_shaders(gpu, seed):
    """scenes.fuzz_shaders: the random state x geometry of the fuzz test drawn with the other parity programs (instancing +
    flat varyings + discard, frag_depth, the early-depth-test programs on every topology, three colour attachments,
    perspective-correct varyings, primitive_index, a depth-only stage).  400 further seeds were run once on the model."""
    from oracle import pyoracle
    from wgpu_cpu_b200.render import SceneRenderer
    dev, queue = gpu
    scene = S.fuzz_shaders(seed)
    ref = pyoracle.render(scene, want_coverage=False)
    r = SceneRenderer(dev, queue, scene)
    r.render()
    got = r.read()
    assert np.array_equal(got.color, ref.color), scene.name
    if ref.depth is not None:
        assert np.array_equal(got.depth.view(np.uint32), ref.depth.view(np.uint32)), scene.name
    for k, (g, e) in enumerate(zip(got.extra_colors, ref.extra_colors), start=1):
        assert np.array_equal(g, e), f"{scene.name}: colour attachment {k}"


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_with_the_opt_in_features(ext, seed):
    """scenes.fuzz_features: the fuzz test's state x geometry on devices with random sets of the opt-in features -- random
    blend states (all thirteen factors, five operations, lines and points included), colour write masks, viewport depth
    ranges -- against the oracle's restatement of them.  300 further seeds were run once on the model."""
    scene = S.fuzz_features(seed)
    got, ref = _render_both(scene, ext)
    assert np.array_equal(got.color, ref.color), (scene.name, scene.features, scene.blend, scene.color_write_mask)
    if ref.depth is not None:
        assert np.array_equal(got.depth.view(np.uint32), ref.depth.view(np.uint32)), scene.name


@pytest.mark.parametrize("seed", range(8))
def test_fuzz_texture_sampling(gpu, seed):
    """scenes.fuzz_textured: texture coordinates far outside [0, 1] and exactly on texel boundaries, every address mode
    per axis, 1 x 1 to odd non-square textures, both sampled formats (binding.rs:93-164).  400 further seeds were run
    once on the model."""
    _compare(S.fuzz_textured(seed), gpu)


@pytest.mark.parametrize("size", [(16384, 40), (40, 16384), (16384, 1)])
def test_maximum_texture_extent(gpu, size):
    """Attachments at the 16384-texel limit: tile column / row 511, the last one the direct bins address."""
    from wgpu_cpu_b200 import api
    _compare(S.random_triangles(size[0], size[1], count=400, seed=5, spread=1.05), gpu)
    with pytest.raises(api.WgpuError):
        gpu[0].create_texture(16385, 8, "rgba8unorm")


def test_several_pipelines_in_one_pass(gpu):
    """A render pass that switches pipelines, bindings, topologies, viewports and scissors between draws (the reference
    replays the sub-commands in order on one State, render_pass/mod.rs:351-388): depth-tested triangles, then lines, then
    an instanced draw that discards, then textured triangles under a scissor -- later draws see the depth and colour of
    earlier ones.  The oracle renders them one after the other, each loading what the previous left."""
    import copy
    from oracle import pyoracle
    from wgpu_cpu_b200.render import SceneRenderer
    dev, queue = gpu
    W, H = 200, 150
    tri = S.random_triangles(W, H, count=120, seed=61, color_format="rgba8unorm")
    lines = S.random_lines(W, H, 80, 62, "line-strip")
    inst = S.features(W, H, instances=2)
    tex = S.fuzz_textured(3)
    tex.width, tex.height, tex.scissor = W, H, (20, 10, 150, 120)
    tex.viewport = (10.0, 5.0, 180.0, 140.0, 0.0, 1.0)
    for s in (lines, inst, tex):
        s.color_format = "rgba8unorm"
    lines.depth_compare, lines.depth_write = "less-equal", False
    scenes = [tri, lines, inst, tex]
    # the oracle: one scene after the other, loading the previous result
    ref = pyoracle.render(scenes[0], want_coverage=False)
    for s in scenes[1:]:
        step = copy.copy(s)
        step.clear_color, step.initial_color = None, ref.color
        step.clear_depth, step.initial_depth = None, ref.depth
        ref = pyoracle.render(step, want_coverage=False)
    # the device: one pass
    rs = [SceneRenderer(dev, queue, s) for s in scenes]
    enc = dev.create_command_encoder()
    with enc.begin_render_pass([{"view": rs[0].target_view, "load": ("clear", tri.clear_color)}],
                               {"view": rs[0].depth_view, "depth_load": ("clear", tri.clear_depth)}) as rp:
        for r in rs:
            r.record_into(rp)
            if r.scene.viewport is None:
                rp.set_viewport(0.0, 0.0, float(W), float(H))          # pass state persists: undo the previous scene's
            if r.scene.scissor is None:
                rp.set_scissor_rect(0, 0, W, H)
    dev.poll(True, queue.submit([enc.finish()]))
    assert np.array_equal(rs[0].target.read(), ref.color)
    assert np.array_equal(rs[0].depth_texture.read().view(np.uint32), ref.depth.view(np.uint32))


def test_recorded_commands_keep_their_resources_alive(gpu):
    """Backend objects are reference counted like the reference's Arc clones (buffer.rs:22-29, command.rs:10-13): a recorded
    pass keeps the pipeline, bind groups, buffers and views it names alive, so the application may drop its handles before
    it submits."""
    import gc
    from oracle import pyoracle
    from wgpu_cpu_b200.render import SceneRenderer
    dev, queue = gpu
    scene = S.hello_texture(160, 110)
    ref = pyoracle.render(scene, want_coverage=False)
    r = SceneRenderer(dev, queue, scene)
    cb = r.encode()
    target, depth = r.target, r.depth_texture
    del r                      # pipeline, module, bind groups, sampler, sampled texture + view, vertex / index / uniform buffers
    gc.collect()
    junk = [dev.create_buffer(1 << 16, mapped_at_creation=False) for _ in range(8)]      # churn the allocator
    dev.poll(True, queue.submit([cb]))
    del junk
    assert np.array_equal(target.read(), ref.color)
    assert np.array_equal(depth.read().view(np.uint32), ref.depth.view(np.uint32))
