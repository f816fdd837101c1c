"""Surface / present (reference surface.rs:51-198) -- the host-side behaviour, without a GPU: capabilities,
configuration checks and the errors the reference panics with, on the compile-only device."""
import ctypes as C

import numpy as np
import pytest

from wgpu_cpu_b200 import api


@pytest.fixture(scope="module")
def offline():
    inst = api.instance()
    adapter = inst.request_adapter()
    dev, queue = adapter.request_device(api.CUDA_DEVICE_COMPILE_ONLY)
    return inst, adapter, dev, queue


def test_capabilities_are_the_references(offline):
    inst, adapter, _, _ = offline
    surface = inst.create_surface()
    assert adapter.is_surface_supported(surface)                        # adapter.rs:44-54
    caps = surface.get_capabilities(adapter)                            # surface.rs:52-72
    assert caps == {"formats": ["bgra8unorm"], "present_modes": ["immediate"], "alpha_modes": ["opaque"],
                    "usages": api.TEXTURE_USAGE["render-attachment"]}


def test_an_unconfigured_surface_has_no_texture(offline):
    inst, _, _, _ = offline
    surface = inst.create_surface()
    for call in (surface.get_current_texture, surface.present):
        with pytest.raises(api.WgpuError) as e:                         # surface.rs:127-130, 174-177: `.expect(..)`
            call()
        assert e.value.status == 1 and "Surface not configured yet" in str(e.value)
    surface.texture_discard()                                           # surface.rs:195-197: nothing


@pytest.mark.parametrize("kwargs, message", [
    (dict(width=64, height=32, format="rgba8unorm"), "Unsupported surface texture format"),             # surface.rs:244-257
    (dict(width=64, height=32, view_formats=["bgra8unorm", "bgra8unorm-srgb"]), "Unsupported surface texture format"),
    (dict(width=0, height=32), "Surface width must not be zero"),                                         # surface.rs:105
    (dict(width=64, height=0), "Surface height must not be zero"),                                        # surface.rs:106
])
def test_configure_refuses_what_the_reference_panics_on(offline, kwargs, message):
    inst, _, dev, _ = offline
    surface = inst.create_surface()
    with pytest.raises(api.WgpuError) as e:
        surface.configure(dev, **kwargs)
    assert e.value.status == 1 and message in str(e.value)
    with pytest.raises(api.WgpuError):
        surface.get_current_texture()


def test_configure_hands_out_one_texture_per_configuration(offline):
    inst, _, dev, _ = offline
    surface = inst.create_surface()
    surface.configure(dev, 64, 32, view_formats=["bgra8unorm"])
    a, b = surface.get_current_texture(), surface.get_current_texture()
    assert a._h.value == b._h.value and (a.width, a.height, a.format, a.status) == (64, 32, "bgra8unorm", 0)   # `configured.buffer.clone()`, SurfaceStatus::Good
    a.create_view()                                                     # usable as an attachment view
    pixels, presents = surface.window_buffer()
    assert pixels.shape == (32, 64, 4) and not pixels.any() and presents == 0
    surface.configure(dev, 16, 8)                                       # a new configuration: a new texture, a new window size
    c = surface.get_current_texture()
    assert c._h.value != a._h.value and (c.width, c.height) == (16, 8)
    assert surface.window_buffer()[0].shape == (8, 16, 4)
    a.create_view()                                                     # the old handle keeps the old texture alive
    with pytest.raises(api.WgpuError) as e:                             # nothing to present without a device to render on
        surface.present()
    assert e.value.status == 4


def test_a_texture_handle_is_not_a_surface(offline):
    inst, adapter, dev, _ = offline
    t = dev.create_texture(4, 4, "bgra8unorm")
    out = C.c_int32(7)
    api._check(api._lib.wgb_adapter_is_surface_supported(adapter._h, t._h, C.byref(out)))
    assert out.value == 0
    assert api._lib.wgb_surface_present(t._h) == 1 and "not a surface" in api._lib.wgb_last_error().decode()
