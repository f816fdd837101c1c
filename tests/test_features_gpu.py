"""pipeline states the reference accepts and ignores, switched on per device with
wgb_device_descriptor.features (parity mode -- no feature bits -- stays bit-exact, tests/test_parity_gpu.py).
The oracle restates these from the WebGPU specification (oracle.h ORC_EXT_*): there is no reference behaviour to
pin them against, so these tests check the CUDA path against that restatement only."""
import numpy as np
import pytest

from wgpu_cpu_b200 import scenes as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ext():
    """Devices by feature set (a device applies exactly the features it was requested with)."""
    from wgpu_cpu_b200 import api
    cache = {}

    def get(features):
        if features not in cache:
            cache[features] = api.instance().request_adapter().request_device(0, features=features)
        return cache[features]
    return get


def _render_both(scene, ext):
    from oracle import pyoracle
    from wgpu_cpu_b200.render import render_scene
    dev, queue = ext(scene.features)
    return render_scene(dev, queue, scene, want_coverage=False), pyoracle.render(scene, want_coverage=False)


def test_viewport_depth_range(ext):
    from wgpu_cpu_b200 import api
    scene = S.random_triangles(count=300, seed=8)
    scene.viewport = (8.0, 4.0, 200.0, 150.0, 0.25, 0.75)
    scene.features = api.FEATURE["VIEWPORT_DEPTH_RANGE"]
    got, ref = _render_both(scene, ext)
    assert np.array_equal(got.depth.view(np.uint32), ref.depth.view(np.uint32))
    assert np.array_equal(got.color, ref.color)
    drawn = ref.depth != np.float32(1.0)
    assert drawn.any()
    # without the feature bit the same scene keeps ndc.z (the reference's behaviour)
    scene.features = 0
    from oracle import pyoracle
    plain = pyoracle.render(scene, want_coverage=False)
    assert not np.array_equal(plain.depth, ref.depth)


@pytest.mark.parametrize("load", [False, True])
@pytest.mark.parametrize("fmt", ["rgba8unorm", "bgra8unorm"])
def test_color_write_mask(ext, load, fmt):
    from wgpu_cpu_b200 import api
    scene = S.random_triangles(count=200, seed=12, color_format=fmt)
    scene.features = api.FEATURE["COLOR_WRITE_MASK"]
    scene.color_write_mask = api.COLOR_WRITE["RED"] | api.COLOR_WRITE["BLUE"]
    scene.clear_color = (0.2, 0.4, 0.6, 1.0)
    if load:
        scene.clear_color = None
        scene.initial_color = np.random.default_rng(2).integers(0, 255, (scene.height, scene.width, 4), dtype=np.uint8)
    got, ref = _render_both(scene, ext)
    assert np.array_equal(got.color, ref.color)
    assert np.array_equal(got.depth.view(np.uint32), ref.depth.view(np.uint32))
    # green and alpha never change
    base = scene.initial_color if load else np.broadcast_to(np.array([51, 102, 153, 255], dtype=np.uint8), ref.color.shape)
    if fmt == "bgra8unorm" and not load:
        base = base[..., [2, 1, 0, 3]]
    assert np.array_equal(ref.color[..., 1], base[..., 1]) and np.array_equal(ref.color[..., 3], base[..., 3])
    assert not np.array_equal(ref.color[..., 0], base[..., 0])


def test_srgb_encode(ext):
    from wgpu_cpu_b200 import api
    scene = S.hello_mesh(200, 160)
    assert scene.color_format.endswith("srgb")
    scene.features = api.FEATURE["SRGB_ENCODE"]
    scene.clear_color = (0.1, 0.2, 0.3, 1.0)
    got, ref = _render_both(scene, ext)
    assert np.array_equal(got.depth.view(np.uint32), ref.depth.view(np.uint32))
    # powf on the device and in libm may differ in the last place: 1 LSB after the 8-bit truncation
    assert np.abs(got.color.astype(np.int32) - ref.color.astype(np.int32)).max() <= 1
    assert tuple(ref.color[0, 0]) == (89, 123, 148, 255)          # OETF of the clear colour, truncated
    scene.features = 0
    from oracle import pyoracle
    assert tuple(pyoracle.render(scene, want_coverage=False).color[0, 0]) == (25, 51, 76, 255)


def test_dynamic_offsets(ext):
    from wgpu_cpu_b200 import api
    scene = S.hello_mesh(160, 120)
    kind, matrix = scene.bindings[(0, 0)]
    junk = np.full(256, 0x7F, dtype=np.uint8)
    scene.bindings[(0, 0)] = (kind, np.concatenate([junk, np.ascontiguousarray(matrix).view(np.uint8).reshape(-1)]))
    scene.features = api.FEATURE["DYNAMIC_OFFSETS"]
    scene.dynamic_bindings = {0: [0]}
    scene.dynamic_offsets = {0: [256]}
    got, ref = _render_both(scene, ext)
    plain = S.hello_mesh(160, 120)
    from oracle import pyoracle
    want = pyoracle.render(plain, want_coverage=False)
    assert np.array_equal(ref.color, want.color)
    assert np.array_equal(got.color, want.color) and np.array_equal(got.depth.view(np.uint32), want.depth.view(np.uint32))


def test_parity_mode_ignores_the_states(ext):
    """A device without feature bits renders what the reference renders, whatever the viewport depth range and the
    write mask say."""
    from oracle import pyoracle
    from wgpu_cpu_b200 import api
    from wgpu_cpu_b200.render import render_scene
    dev, queue = api.instance().request_adapter().request_device(0)
    scene = S.random_triangles(count=150, seed=4)
    scene.viewport = (0.0, 0.0, float(scene.width), float(scene.height), 0.3, 0.6)
    scene.color_write_mask = 1
    ref = pyoracle.render(scene, want_coverage=False)
    got = render_scene(dev, queue, scene, want_coverage=False)
    assert np.array_equal(got.color, ref.color) and np.array_equal(got.depth.view(np.uint32), ref.depth.view(np.uint32))


BLENDS = {
    "alpha": {"color": ("src-alpha", "one-minus-src-alpha", "add"), "alpha": ("one", "one-minus-src-alpha", "add")},
    "additive": {"color": ("one", "one", "add"), "alpha": ("one", "one", "add")},
    "multiply_revsub": {"color": ("dst", "zero", "add"), "alpha": ("one", "one", "reverse-subtract")},
    "min_max": {"color": ("one", "one", "min"), "alpha": ("one", "one", "max")},
    "constant_saturate": {"color": ("constant", "one-minus-constant", "add"), "alpha": ("src-alpha-saturated", "dst-alpha", "subtract")},
}


def _translucent(scene, seed):
    """random alpha per vertex (the hello_mesh shader passes the vertex colour through)"""
    v = scene.vertex_buffers[0].view(np.float32).reshape(-1, 8).copy()
    v[:, 7] = np.random.default_rng(seed).random(v.shape[0], dtype=np.float32)
    scene.vertex_buffers[0] = v.view(np.uint8).reshape(-1)


@pytest.mark.parametrize("name", sorted(BLENDS))
@pytest.mark.parametrize("depth", [None, "less"])
def test_blending(ext, name, depth):
    """WGB_FEATURE_BLEND: fragments blend against the stored 8-bit texel in primitive order (ordered tile kernel)."""
    from wgpu_cpu_b200 import api
    scene = S.random_triangles(count=220, seed=5, color_format="rgba8unorm", depth_compare=depth, depth_write=depth is not None)
    _translucent(scene, 9)
    scene.features = api.FEATURE["BLEND"]
    scene.blend = BLENDS[name]
    scene.blend_constant = (0.25, 0.5, 0.75, 0.4)
    scene.clear_color = (0.1, 0.3, 0.2, 0.5)
    got, ref = _render_both(scene, ext)
    assert np.array_equal(got.color, ref.color)
    if depth is not None:
        assert np.array_equal(got.depth.view(np.uint32), ref.depth.view(np.uint32))
    # blending really happened: the frame differs from the unblended one
    scene.features = 0
    from oracle import pyoracle
    assert not np.array_equal(pyoracle.render(scene, want_coverage=False).color, ref.color)



def test_blending_with_srgb_target_mask_and_indexed_mesh(ext):
    from wgpu_cpu_b200 import api
    scene = S.synthetic_grid(320, 200, n=40, layers=3)
    _translucent(scene, 3)
    scene.features = api.FEATURE["BLEND"] | api.FEATURE["SRGB_ENCODE"] | api.FEATURE["COLOR_WRITE_MASK"]
    scene.blend = BLENDS["alpha"]
    scene.color_write_mask = 7
    scene.clear_color = (0.2, 0.2, 0.2, 1.0)
    got, ref = _render_both(scene, ext)
    assert np.array_equal(got.depth.view(np.uint32), ref.depth.view(np.uint32))
    # sRGB decode / encode go through powf on both sides: 1 LSB per blend step can compound over the layers
    assert np.abs(got.color.astype(np.int32) - ref.color.astype(np.int32)).max() <= 3
    assert (got.color == ref.color).mean() > 0.99
