"""Parity of the B200 draw path (through the C ABI) against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): coverage masks and depth-test outcomes bit-exact, shaded colour
within 1 LSB for unorm8 targets -- the tests below demand byte-exact colour as well, since both
sides run the same IEEE operation sequence."""
import numpy as np
import pytest

from wgpu_cpu_b200 import scenes as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from wgpu_cpu_b200 import api
    dev, queue = api.instance().request_adapter().request_device(0)
    return dev, queue


def _compare(scene, gpu, use_emitted=False):
    from oracle import pyoracle
    from wgpu_cpu_b200.render import render_scene
    dev, queue = gpu
    ref = pyoracle.render(scene)
    got = render_scene(dev, queue, scene, want_coverage=True, use_emitted=use_emitted)
    msg = []
    if not np.array_equal(got.coverage, ref.coverage):
        bad = np.argwhere(got.coverage != ref.coverage)
        msg.append(f"coverage differs at {len(bad)} pixels, first (y,x)={bad[:5].tolist()} "
                   f"got={got.coverage[tuple(bad[0])]} ref={ref.coverage[tuple(bad[0])]}")
    if ref.depth is not None:
        gd, rd = got.depth.view(np.uint32), ref.depth.view(np.uint32)
        if not np.array_equal(gd, rd):
            bad = np.argwhere(gd != rd)
            msg.append(f"depth differs at {len(bad)} pixels, first (y,x)={bad[:5].tolist()} "
                       f"got={got.depth[tuple(bad[0])]!r} ref={ref.depth[tuple(bad[0])]!r}")
    if not np.array_equal(got.color, ref.color):
        bad = np.argwhere((got.color != ref.color).any(axis=2))
        msg.append(f"colour differs at {len(bad)} pixels, first (y,x)={bad[:5].tolist()} "
                   f"got={got.color[tuple(bad[0])].tolist()} ref={ref.color[tuple(bad[0])].tolist()}")
    assert not msg, f"{scene.name}: " + "; ".join(msg)
    # (the oracle counts stage invocations: fewer than the rasterised fragments once an early depth test rejects some)
    assert got.stats["fragments"] == ref.stats["fragments_shaded"] or scene.shader in ("features", "early_force", "early_allow", "depth_only"), \
        f"{scene.name}: fragment count {got.stats['fragments']} != {ref.stats['fragments_shaded']}"
    # the production configuration: no coverage capture, so the hierarchical depth test is active
    fast = render_scene(dev, queue, scene, want_coverage=False, use_emitted=use_emitted)
    assert np.array_equal(fast.color, ref.color), f"{scene.name}: colour differs with the hierarchical depth test on"
    if ref.depth is not None:
        assert np.array_equal(fast.depth.view(np.uint32), ref.depth.view(np.uint32)), \
            f"{scene.name}: depth differs with the hierarchical depth test on"
    return got, ref


@pytest.mark.parametrize("variant", ["default", "cull_front", "draw_backwards", "draw_backwards_no_cull", "lines"])
def test_colored_triangle(gpu, variant):
    _compare(S.colored_triangle(variant), gpu)


def test_colored_triangle_golden_identities(gpu):
    """tests/reference/*.png are LFS pointers; their oids still pin A == D and B == C == clear."""
    from wgpu_cpu_b200.render import render_scene
    dev, queue = gpu
    img = {v: render_scene(dev, queue, S.colored_triangle(v), use_emitted=False).color
           for v in ["default", "cull_front", "draw_backwards", "draw_backwards_no_cull"]}
    assert np.array_equal(img["default"], img["draw_backwards_no_cull"])
    assert np.array_equal(img["cull_front"], img["draw_backwards"])
    assert (img["cull_front"][..., :3] == 0).all() and (img["cull_front"][..., 3] == 255).all()
    assert img["default"][..., :3].any()


def test_hello_mesh_c1(gpu):
    _compare(S.hello_mesh(512, 512), gpu)


def test_hello_texture_c2_small(gpu):
    _compare(S.hello_texture(640, 360), gpu)


def test_synthetic_c3_small(gpu):
    _compare(S.synthetic_grid(384, 216, n=113, layers=4), gpu)


def test_procedural_c4_small(gpu):
    _compare(S.procedural(480, 270), gpu)


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_random_triangles_clipped(gpu, seed):
    got, ref = _compare(S.random_triangles(seed=seed), gpu)
    assert got.stats["clipped_primitives"] > 0


@pytest.mark.parametrize("compare,write", [("less", True), ("less-equal", True), ("greater", True), ("greater-equal", True),
                                           ("always", True), ("always", False), ("never", True), ("less", False),
                                           ("equal", True), ("not-equal", False), ("greater", False)])
def test_depth_modes(gpu, compare, write):
    s = S.random_triangles(count=150, seed=41, depth_compare=compare, depth_write=write,
                           clear_depth=0.5 if compare in ("greater", "greater-equal", "equal", "not-equal") else 1.0)
    s.name += f"_{compare}_{int(write)}"
    _compare(s, gpu)


def test_no_depth_attachment_last_wins(gpu):
    s = S.random_triangles(count=200, seed=43, has_depth=False, clear_depth=None, depth_compare=None, depth_write=False)
    _compare(s, gpu)


@pytest.mark.parametrize("index_format,restart", [("uint16", True), ("uint32", True), ("uint16", False)])
def test_triangle_strip(gpu, index_format, restart):
    _compare(S.quad_strip(index_format=index_format, restart=restart), gpu)


@pytest.mark.parametrize("topology", ["line-list", "line-strip"])
def test_lines(gpu, topology):
    _compare(S.random_lines(topology=topology), gpu)


def test_points(gpu):
    _compare(S.random_points(), gpu)


def test_features_instancing_flat_discard(gpu):
    _compare(S.features(), gpu)


def test_frag_depth(gpu):
    _compare(S.frag_depth(), gpu)




def test_viewport_and_scissor(gpu):
    s = S.random_triangles(count=120, seed=47)
    s.viewport = (20.0, 10.0, 200.0, 150.0, 0.0, 1.0)
    s.scissor = (40, 30, 120, 90)
    s.name += "_vp_sc"
    _compare(s, gpu)


def test_load_op_load(gpu):
    rng = np.random.default_rng(3)
    s = S.random_triangles(count=100, seed=49, clear_color=None, clear_depth=None)
    s.initial_color = rng.integers(0, 255, (s.height, s.width, 4), dtype=np.uint8)
    s.initial_depth = rng.random((s.height, s.width), dtype=np.float32)
    s.name += "_load"
    _compare(s, gpu)


def test_every_texel_value_decodes_as_the_division_does(gpu):
    """u8 as f32 / 255.0 (texture.rs:170-188) is computed without a division on the device (wgb_unorm8); a texture of
    random bytes -- every value in every channel -- sampled all over the bunny must give the oracle's bytes: the target's
    truncating encode turns a last-bit difference of the quotient into a different byte."""
    rng = np.random.default_rng(255)
    s = S.hello_texture(320, 180)
    s.bindings = dict(s.bindings)
    for key, res in list(s.bindings.items()):
        if res[0] == "texture":
            img = rng.integers(0, 256, res[1].shape, dtype=np.uint8)
            img.reshape(-1)[:1024] = np.repeat(np.arange(256, dtype=np.uint8), 4)
            s.bindings[key] = ("texture", img) + tuple(res[2:])
    s.name += "_random_texels"
    _compare(s, gpu)


@pytest.mark.parametrize("size", [(200, 100), (36, 36), (4, 4), (64, 33), (132, 64)])
@pytest.mark.parametrize("load", [False, True])
def test_tensor_map_tiles_on_the_edges(gpu, size, load):
    """Widths that are a multiple of four texels take the tensor-map path of the write-back (one cp.async.bulk.tensor.2d
    per tile and attachment): tiles cut by the right and bottom edges store only the texels inside the attachment, and
    with LoadOp::Load the texels of the box outside it arrive as zeros without disturbing the byte count the mbarrier
    waits for."""
    rng = np.random.default_rng(size[0] * 1000 + size[1])
    kw = dict(clear_color=None, clear_depth=None) if load else {}
    s = S.random_triangles(size[0], size[1], count=60, seed=5 + size[0], **kw)
    if load:
        s.initial_color = rng.integers(0, 255, (s.height, s.width, 4), dtype=np.uint8)
        s.initial_depth = rng.random((s.height, s.width), dtype=np.float32)
    s.name += f"_tmap_{int(load)}"
    _compare(s, gpu)


@pytest.mark.parametrize("scene_name", ["hello_mesh", "synthetic_clipped", "random_clipped"])
@pytest.mark.parametrize("bands", [2, 3, 8])
def test_sort_first_bands_reassemble_the_single_gpu_frame(gpu, bands, scene_name):
    """each band renders with the full scene and only its tile rows; per-pixel primitive order
    is preserved inside a band, so the reassembled frame is bit-identical to the oracle's."""
    from oracle import pyoracle
    from wgpu_cpu_b200 import api
    from wgpu_cpu_b200.multigpu import band_rows
    from wgpu_cpu_b200.render import render_scene
    # synthetic_clipped: an indexed mesh whose border triangles cross the clip volume (the cached geometry kernel
    # drops primitives -- clipped ones included -- whose vertex rows miss the band); random_clipped: non-indexed
    scene = {"hello_mesh": lambda: S.hello_mesh(320, 200),
             "synthetic_clipped": lambda: S.synthetic_grid(320, 200, n=60, layers=2),
             "random_clipped": lambda: S.random_triangles(320, 200, count=300, seed=3)}[scene_name]()
    ref = pyoracle.render(scene)
    color = np.zeros_like(ref.color)
    depth = np.zeros_like(ref.depth)
    frags = 0
    for r in range(bands):
        dev, queue = api.instance().request_adapter().request_device(0, band_rank=r, band_count=bands)
        got = render_scene(dev, queue, scene, want_coverage=False)
        a, b = band_rows(scene.height, r, bands)
        assert (a, b) == dev.band_rows(scene.height)
        color[a:b] = got.color[a:b]
        depth[a:b] = got.depth[a:b]
        # rows outside the band are never touched (textures start zeroed)
        assert not got.color[:a].any() and not got.color[b:].any()
        # every fragment is rasterised by exactly one band (counted with the hierarchical depth test off)
        frags += render_scene(dev, queue, scene, want_coverage=True).stats["fragments"]
    assert np.array_equal(color, ref.color)
    assert np.array_equal(depth.view(np.uint32), ref.depth.view(np.uint32))
    assert frags == ref.stats["fragments_shaded"]


def test_readback_through_copy_texture_to_buffer(gpu):
    """the standard wgpu read-back (copy_texture_to_buffer with a 256-byte row pitch, then
    map_async(Read)) returns the same texels as the direct texture read the reference's tests use."""
    from wgpu_cpu_b200 import api
    from wgpu_cpu_b200.render import SceneRenderer
    dev, queue = gpu
    scene = S.hello_mesh(200, 120)
    r = SceneRenderer(dev, queue, scene)
    pitch = (scene.width * 4 + 255) // 256 * 256
    staging = dev.create_buffer(pitch * scene.height, api.BUFFER_USAGE["MAP_READ"] | api.BUFFER_USAGE["COPY_DST"])
    enc = dev.create_command_encoder()
    enc.clear_buffer(staging)
    cb_render = r.encode()
    enc.copy_texture_to_buffer(r.target, staging, bytes_per_row=pitch)
    idx = queue.submit([cb_render, enc.finish()])
    dev.poll(True, idx)
    staging.map_async("read")
    rows = staging.get_mapped_range().reshape(scene.height, pitch)[:, :scene.width * 4].reshape(scene.height, scene.width, 4).copy()
    staging.unmap()
    assert np.array_equal(rows, r.target.read())
    assert rows[..., :3].any()


def test_encoder_copies_roundtrip(gpu):
    from wgpu_cpu_b200 import api
    dev, queue = gpu
    rng = np.random.default_rng(5)
    img = rng.integers(0, 255, (48, 64, 4), dtype=np.uint8)
    src = dev.create_buffer_init(img, api.BUFFER_USAGE["COPY_SRC"])
    a = dev.create_texture(64, 48, "rgba8unorm")
    b = dev.create_texture(80, 60, "rgba8unorm")
    out = dev.create_buffer(64 * 48 * 4, api.BUFFER_USAGE["COPY_DST"] | api.BUFFER_USAGE["MAP_READ"])
    out2 = dev.create_buffer(64 * 48 * 4, api.BUFFER_USAGE["COPY_DST"] | api.BUFFER_USAGE["MAP_READ"])
    enc = dev.create_command_encoder()
    enc.copy_buffer_to_texture(src, a)
    enc.copy_texture_to_texture(a, b, size=(32, 20), src_origin=(8, 4), dst_origin=(40, 30))
    enc.copy_texture_to_buffer(a, out)
    enc.copy_buffer_to_buffer(out, 256, out2, 512, 1024)
    idx = queue.submit([enc.finish()])
    dev.poll(True, idx)
    assert np.array_equal(a.read(), img)
    tb = b.read()
    assert np.array_equal(tb[30:50, 40:72], img[4:24, 8:40]) and not tb[:30].any() and not tb[:, :40].any()
    out.map_async("read")
    assert np.array_equal(out.get_mapped_range(), img.reshape(-1))
    out.unmap()
    out2.map_async("read")
    got = out2.get_mapped_range().copy()
    out2.unmap()
    assert np.array_equal(got[512:1536], img.reshape(-1)[256:1280]) and not got[:512].any() and not got[1536:].any()
    enc = dev.create_command_encoder()
    enc.clear_texture(a)
    idx = queue.submit([enc.finish()])
    dev.poll(True, idx)
    assert not a.read().any()


def test_multiple_draws_ragged_and_empty(gpu):
    _compare(S.multi_draw(), gpu)


def test_instance_step_mode_buffer(gpu):
    _compare(S.instanced_step_mode(), gpu)


def test_huge_triangles_and_slivers(gpu):
    got, ref = _compare(S.huge_triangles(), gpu)
    assert got.stats["big_primitives"] > 0 and got.stats["clipped_primitives"] > 0


@pytest.mark.parametrize("size", [(333, 77), (31, 33), (1, 1), (65, 1)])
def test_odd_framebuffer_sizes(gpu, size):
    _compare(S.odd_sizes(*size), gpu)


def test_empty_scissor_draws_nothing_but_clears(gpu):
    s = S.random_triangles(count=50, seed=91)
    s.scissor = (10, 10, 0, 0)
    s.name += "_empty_scissor"
    got, ref = _compare(s, gpu)
    assert got.stats["fragments"] == 0 and (got.color[..., 3] == 255).all() and not got.color[..., :3].any()


def test_two_passes_in_one_submission_second_loads(gpu):
    """A cleared pass followed by a LoadOp::Load pass on the same attachments (state.rs:135-145)."""
    from oracle import pyoracle
    from wgpu_cpu_b200.render import SceneRenderer
    dev, queue = gpu
    s1 = S.random_triangles(count=80, seed=101)
    r1 = SceneRenderer(dev, queue, s1)
    cb1 = r1.encode()
    s2 = S.random_triangles(count=80, seed=102, clear_color=None, clear_depth=None)
    r2 = SceneRenderer(dev, queue, s2, target=r1.target)
    r2.depth_texture, r2.depth_view = r1.depth_texture, r1.depth_view
    cb2 = r2.encode()
    idx = queue.submit([cb1, cb2])
    dev.poll(True, idx)
    f1 = pyoracle.render(s1)
    s2.initial_color, s2.initial_depth = f1.color, f1.depth
    f2 = pyoracle.render(s2)
    assert np.array_equal(r1.target.read(), f2.color)
    assert np.array_equal(r1.depth_texture.read().view(np.uint32), f2.depth.view(np.uint32))


def test_out_of_bounds_index_is_an_error_and_leaves_the_target_untouched(gpu):
    """The reference panics on the slice index (index.rs:45-51); here the pass aborts before any attachment write
    and the error surfaces at poll."""
    from wgpu_cpu_b200 import api
    from wgpu_cpu_b200.render import SceneRenderer
    dev, queue = gpu
    s = S.multi_draw()
    s.draws = [S.Draw(True, 0, 90, 400)]          # base_vertex pushes every fetch outside the vertex buffer
    r = SceneRenderer(dev, queue, s)
    idx = r.submit()
    with pytest.raises(api.WgpuError) as e:
        dev.poll(True, idx)
    assert e.value.status == 6
    assert not r.target.read().any()


def test_index_range_outside_the_index_buffer_is_an_error(gpu):
    from wgpu_cpu_b200 import api
    from wgpu_cpu_b200.render import SceneRenderer
    dev, queue = gpu
    s = S.multi_draw()
    s.draws = [S.Draw(True, 100, 200)]            # the index buffer holds 180 indices
    r = SceneRenderer(dev, queue, s)
    idx = r.submit()
    with pytest.raises(api.WgpuError) as e:
        dev.poll(True, idx)
    assert e.value.status == 6 and "index buffer" in str(e.value)
    assert not r.target.read().any()


def test_pinned_uploads_on_the_copy_stream_are_ordered_with_rendering(gpu):
    """wgb_queue_write_buffer from pinned memory (>= 1 MiB) runs on the device's copy stream; each frame must see
    exactly the data written before its submission (write-after-read and read-after-write hazards per buffer)."""
    import torch
    from oracle import pyoracle
    from wgpu_cpu_b200.render import SceneRenderer
    dev, queue = gpu
    scn = [S.random_triangles(count=12000, seed=sd, spread=1.2) for sd in (21, 22)]
    assert scn[0].vertex_buffers[0].nbytes >= (1 << 20)
    refs = [pyoracle.render(s, want_coverage=False).color for s in scn]
    pinned = []
    for s in scn:
        t = torch.empty(s.vertex_buffers[0].nbytes, dtype=torch.uint8, pin_memory=True)
        t.numpy()[:] = s.vertex_buffers[0]
        pinned.append(t)
    r = SceneRenderer(dev, queue, scn[0])
    for k in (1, 0, 1, 1, 0):
        queue.write_buffer(r.vertex_buffers[0], 0, pinned[k].numpy())
        r.render()
        # the next upload is enqueued before this frame is read back
        queue.write_buffer(r.vertex_buffers[0], 0, pinned[1 - k].numpy())
        queue.write_buffer(r.vertex_buffers[0], 0, pinned[k].numpy())
        assert np.array_equal(r.target.read(), refs[k])
        r.render()
        assert np.array_equal(r.target.read(), refs[k])
    # a small (render-stream) write after a pending copy-stream write must land after it
    queue.write_buffer(r.vertex_buffers[0], 0, pinned[0].numpy())
    queue.write_buffer(r.vertex_buffers[0], 0, scn[1].vertex_buffers[0][:4096])
    r.render()
    mixed = S.random_triangles(count=12000, seed=21, spread=1.2)
    mixed.vertex_buffers[0][:4096] = scn[1].vertex_buffers[0][:4096]
    assert np.array_equal(r.target.read(), pyoracle.render(mixed, want_coverage=False).color)


def test_direct_bins_overflow_replays_then_learns_the_capacity(gpu):
    """Direct binning gives every tile a fixed number of bin slots; a scene that piles its primitives into a few
    tiles overflows them on first sight, the pass is aborted before any texel is written, and the draw replays with
    the capacity the tile kernel reported."""
    from wgpu_cpu_b200.render import SceneRenderer
    from oracle import pyoracle
    dev, queue = gpu
    scene = S.random_triangles(count=4000, seed=31, spread=0.12)
    ref = pyoracle.render(scene, want_coverage=False)
    r = SceneRenderer(dev, queue, scene)
    st = r.render()
    assert st["replays"] >= 1
    f = r.read()
    assert np.array_equal(f.color, ref.color) and np.array_equal(f.depth.view(np.uint32), ref.depth.view(np.uint32))
    st = r.render()
    assert st["replays"] == 0
    assert np.array_equal(r.read().color, ref.color)


def test_count_scan_fill_binning_matches_direct_binning(monkeypatch):
    """WGB_NO_DIRECT_BINS=1 keeps the exact-size bins (count, scan, fill kernels); both binning modes must give the
    same frame."""
    from wgpu_cpu_b200 import api
    from wgpu_cpu_b200.render import render_scene
    monkeypatch.setenv("WGB_NO_DIRECT_BINS", "1")
    dev2, queue2 = api.instance().request_adapter().request_device(0)
    monkeypatch.delenv("WGB_NO_DIRECT_BINS")
    dev1, queue1 = api.instance().request_adapter().request_device(0)
    for scene in (S.synthetic_grid(640, 360, n=120, layers=3), S.random_triangles(count=900, seed=5), S.hello_mesh(256, 256)):
        a = render_scene(dev1, queue1, scene)
        b = render_scene(dev2, queue2, scene)
        assert a.stats["kernel_launches"] < b.stats["kernel_launches"]
        assert np.array_equal(a.color, b.color) and np.array_equal(a.coverage, b.coverage)
        assert np.array_equal(a.depth.view(np.uint32), b.depth.view(np.uint32))
        assert a.stats["bin_pairs"] == b.stats["bin_pairs"] and a.stats["fragments"] == b.stats["fragments"]





@pytest.mark.parametrize("order", ["front_to_back", "back_to_front"])
def test_hierarchical_depth_test_drops_hidden_triangles_without_changing_the_frame(gpu, order):
    """Layered small triangles (the C3 structure): with the near layer drawn first most of the far layers is dropped
    before rasterisation; the frame must not change, whatever the order."""
    from oracle import pyoracle
    from wgpu_cpu_b200.render import render_scene
    dev, queue = gpu
    scene = S.synthetic_grid(960, 540, n=280, layers=4)
    if order == "back_to_front":
        idx = scene.index_data.reshape(4, -1)
        scene.index_data = np.ascontiguousarray(idx[::-1]).reshape(-1)
    ref = pyoracle.render(scene, want_coverage=False)
    got = render_scene(dev, queue, scene, want_coverage=False)
    assert np.array_equal(got.color, ref.color)
    assert np.array_equal(got.depth.view(np.uint32), ref.depth.view(np.uint32))
    if order == "front_to_back":
        assert got.stats["hiz_culled"] > 0.25 * got.stats["bin_pairs"]
    full = render_scene(dev, queue, scene, want_coverage=True)
    assert full.stats["hiz_culled"] == 0 and full.stats["fragments"] == ref.stats["fragments_shaded"]
    assert got.stats["fragments"] <= full.stats["fragments"]


def test_dump_texture_png(gpu, tmp_path):
    """dump_texture (lib.rs:111-158) of the colour target and of the depth attachment."""
    from tests.test_c_abi import _decode_png
    from wgpu_cpu_b200 import api
    from wgpu_cpu_b200.render import SceneRenderer
    dev, queue = gpu
    r = SceneRenderer(dev, queue, S.hello_mesh(96, 64))
    r.render()
    api.dump_texture(r.target, str(tmp_path / "color.png"))
    assert np.array_equal(_decode_png(str(tmp_path / "color.png")), r.target.read())
    r.depth_texture.dump_png(str(tmp_path / "depth.png"))
    d = r.depth_texture.read()
    want = np.clip(np.trunc(d * np.float32(255.0)), 0, 255).astype(np.uint8)
    assert np.array_equal(_decode_png(str(tmp_path / "depth.png"))[:, :, 0], want)


@pytest.mark.parametrize("seed", range(64))
def test_fuzz(gpu, seed):
    """Random pipeline state x random geometry (scenes.fuzz): framebuffer sizes from 1x1, all topologies, u16/u32
    indices with base vertex / first index / primitive restart, cull and depth states, viewports reaching outside
    the framebuffer, scissors, load ops, instancing -- coverage, depth and colour bit-exact, with and without the
    hierarchical depth test."""
    _compare(S.fuzz(seed), gpu)


@pytest.mark.parametrize("scene_name", ["random", "indexed_small", "huge"])
def test_not_equal_with_depth_write_takes_the_ordered_kernel(gpu, scene_name):
    """NotEqual + depth write: every fragment's outcome depends on the one before it, so there is no closed form; the
    ordered tile kernel applies the fragments of each pixel in primitive order, like the reference's loop."""
    scene = {"random": lambda: S.random_triangles(count=250, seed=77, clear_depth=0.5),
             "indexed_small": lambda: S.synthetic_grid(300, 200, n=50, layers=3),
             "huge": lambda: S.huge_triangles()}[scene_name]()
    scene.depth_compare, scene.depth_write = "not-equal", True
    scene.name += "_not_equal_write"
    _compare(scene, gpu)


@pytest.mark.parametrize("fmt", ["rgba8unorm", "rgba8unorm-srgb", "bgra8unorm", "bgra8unorm-srgb", "r8unorm", "rg8unorm", "rgba8snorm"])
@pytest.mark.parametrize("size", [(256, 192), (250, 190)])
def test_color_target_formats(gpu, fmt, size):
    """TexelWriter::from_color (texture.rs:373-411): every colour target format the reference can write, incl. the
    1- and 2-byte ones (per-texel stores instead of the staged TMA write-back) and Rgba8Snorm, which the reference
    encodes like unorm; widths that are and are not a multiple of 4 texels."""
    _compare(S.random_triangles(size[0], size[1], count=200, seed=17, color_format=fmt), gpu)


@pytest.mark.parametrize("addr", [("clamp-to-edge", "clamp-to-edge"), ("repeat", "mirror-repeat"), ("mirror-repeat", "clamp-to-edge"),
                                  ("mirror-repeat", "repeat")])
def test_sampler_address_modes(gpu, addr):
    """texel_coordinate (binding.rs:151-164): ClampToEdge / Repeat / MirrorRepeat per axis, uv up to 5 (hello_texture.rs:651-659)."""
    scene = S.hello_texture(320, 200)
    key = [k for k, v in scene.bindings.items() if v[0] == "sampler"][0]
    scene.bindings[key] = ("sampler", addr[0], addr[1])
    scene.name += f"_{addr[0]}_{addr[1]}"
    _compare(scene, gpu)
