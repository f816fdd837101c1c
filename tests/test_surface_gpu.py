"""Surface / present on the device (reference surface.rs:116-193): a frame rendered into the surface's texture reaches the
window -- the host pixel sink -- byte for byte as the oracle renders it into a Bgra8Unorm target, through the present
callback and in the window buffer; frames in flight are waited for by present, as the reference waits for the texture's
write guard."""
import numpy as np
import pytest

from oracle import pyoracle
from wgpu_cpu_b200 import api, scenes
from wgpu_cpu_b200.render import SceneRenderer

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    inst = api.instance()
    adapter = inst.request_adapter()
    dev, queue = adapter.request_device(0)
    return inst, adapter, dev, queue


def _bgra(scene):
    scene.color_format = "bgra8unorm"
    return scene


def test_present_delivers_the_oracles_frame(ctx):
    inst, adapter, dev, queue = ctx
    seen = []
    surface = inst.create_surface(lambda pixels: seen.append(pixels.copy()))
    assert adapter.is_surface_supported(surface)
    scene = _bgra(scenes.hello_mesh(160, 120))
    ref = pyoracle.render(scene)
    surface.configure(dev, scene.width, scene.height)
    frame = surface.get_current_texture()
    r = SceneRenderer(dev, queue, scene, target=frame)
    r.submit()                                   # not waited for: present does
    surface.present()
    window, presents = surface.window_buffer()
    assert presents == 1 and len(seen) == 1
    assert np.array_equal(seen[0], ref.color) and np.array_equal(window, ref.color)
    assert np.array_equal(frame.read(), ref.color)            # `target.copy_from_slice(&*source)`: the texture's bytes as they are


def test_every_present_shows_the_frame_submitted_before_it(ctx):
    inst, _, dev, queue = ctx
    digests = []
    surface = inst.create_surface(lambda pixels: digests.append(pixels.copy()))
    w, h = 96, 64
    surface.configure(dev, w, h)
    want = []
    for k, scene in enumerate((scenes.colored_triangle("default", w, h), scenes.hello_mesh(w, h), scenes.colored_triangle("default", w, h))):
        scene = _bgra(scene)
        want.append(pyoracle.render(scene).color)
        SceneRenderer(dev, queue, scene, target=surface.get_current_texture()).submit()
        surface.present()
        assert surface.window_buffer()[1] == k + 1
    assert len(digests) == 3 and all(np.array_equal(a, b) for a, b in zip(digests, want))
    assert not np.array_equal(want[0], want[1])


def test_reconfigure_resizes_the_window(ctx):
    inst, _, dev, queue = ctx
    surface = inst.create_surface()
    for w, h in ((64, 48), (130, 70)):          # (a width that is not a multiple of the tile size)
        scene = _bgra(scenes.colored_triangle("default", w, h))
        surface.configure(dev, w, h)
        assert not surface.window_buffer()[0].any()            # Texture::new: zeroed (surface.rs:94-101)
        SceneRenderer(dev, queue, scene, target=surface.get_current_texture()).render()
        surface.present()
        assert np.array_equal(surface.window_buffer()[0], pyoracle.render(scene).color)


def test_a_multisample_state_is_carried_and_not_applied(ctx):
    """SURVEY 8 f.4, second half: the reference keeps `MultisampleState` in the pipeline (pipeline.rs:67,98) and has the
    WebGPU sample positions in a table (raster.rs:45-63), but its rasteriser emits `sample_index: None` for every fragment
    (raster.rs:213,259,351) -- one sample at the pixel -- so a count of 4 renders the single-sample frame.  So does this."""
    _, _, dev, queue = ctx
    scene = scenes.hello_mesh(96, 64)
    ref = pyoracle.render(scene)
    scene.multisample_count = 4
    r = SceneRenderer(dev, queue, scene)
    r.render()
    f = r.read()
    assert np.array_equal(f.color, ref.color) and np.array_equal(f.depth.view(np.uint32), ref.depth.view(np.uint32))


def test_the_present_callback_may_ask_for_the_next_texture(ctx):
    """(A window loop that requests its next frame from inside the present notification must not deadlock.)"""
    inst, _, dev, queue = ctx
    nxt = []
    surface = inst.create_surface(lambda pixels: nxt.append(surface.get_current_texture()))
    scene = _bgra(scenes.colored_triangle("default", 64, 48))
    surface.configure(dev, 64, 48)
    SceneRenderer(dev, queue, scene, target=surface.get_current_texture()).submit()
    surface.present()
    assert len(nxt) == 1 and (nxt[0].width, nxt[0].height) == (64, 48)
    assert np.array_equal(nxt[0].read(), pyoracle.render(scene).color)
