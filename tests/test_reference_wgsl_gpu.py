"""The reference's own shader text, unmodified (tests/golden/reference_wgsl/: byte-for-byte copies of
wgpu-cpu/examples/hello_mesh.wgsl, hello_texture.wgsl, hello_shader.wgsl and wgpu-cpu-tests/src/tests/colored_triangle.wgsl
-- test fixtures: "an unmodified wgpu app" hands the backend exactly this text).  The scenes elsewhere in the suite
use this repository's own restatement of the same programs; here the reference's files go through the WGSL -> CUDA
emitter and NVRTC and the frames are compared with the oracle."""
import os

import numpy as np
import pytest

from wgpu_cpu_b200 import scenes as S

DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_wgsl")
FILES = ("hello_mesh", "hello_texture", "hello_shader", "colored_triangle")
def text(name):
    with open(os.path.join(DIR, name + ".wgsl")) as f:
        return f.read()


@pytest.mark.parametrize("name", FILES)
def test_the_reference_files_translate(name):
    """CPU tier: both entry points of every file go through the emitter."""
    from wgpu_cpu_b200 import api
    src = text(name)
    vs = api.translate_wgsl(src, api.STAGE_VERTEX, "vs_main")
    fs = api.translate_wgsl(src, api.STAGE_FRAGMENT, "fs_main")
    assert "wgb_vs_entry" in vs and "wgb_fs_entry" in fs
    # @interpolate(linear, sample): linear interpolation, no perspective correction (fragment.rs:346-368)
    assert "#define WGB_FS_WRITES_FRAG_DEPTH 0" in fs and "#define WGB_FS_MAY_DISCARD 0" in fs


def test_the_copies_are_the_reference_files():
    """Where the reference checkout is present (the build container), the fixtures are byte-identical to it."""
    ref = {"hello_mesh": "/root/reference/wgpu-cpu/examples/hello_mesh.wgsl", "hello_texture": "/root/reference/wgpu-cpu/examples/hello_texture.wgsl",
           "hello_shader": "/root/reference/wgpu-cpu/examples/hello_shader.wgsl",
           "colored_triangle": "/root/reference/wgpu-cpu-tests/src/tests/colored_triangle.wgsl"}
    if not all(os.path.exists(p) for p in ref.values()):
        pytest.skip("no reference checkout on this machine")
    for name, p in ref.items():
        assert open(p, "rb").read() == open(os.path.join(DIR, name + ".wgsl"), "rb").read(), name


@pytest.mark.gpu
@pytest.mark.parametrize("name,scene", [
    ("hello_mesh", lambda: S.hello_mesh(256, 256)),
    ("hello_texture", lambda: S.hello_texture(320, 180)),
    ("hello_shader", lambda: S.colored_triangle("default", 128, 128)),
    ("colored_triangle", lambda: S.colored_triangle("default", 128, 128)),
    ("colored_triangle", lambda: S.colored_triangle("lines", 128, 128)),
    ("hello_mesh", lambda: S.synthetic_grid(384, 216, n=113, layers=4)),
])
def test_the_reference_text_renders_the_oracle_frame(name, scene):
    from oracle import pyoracle
    from wgpu_cpu_b200 import api
    from wgpu_cpu_b200.render import render_scene
    dev, queue = api.instance().request_adapter().request_device(0)
    sc = scene()
    ref = pyoracle.render(sc)
    got = render_scene(dev, queue, sc, want_coverage=True, wgsl=text(name))
    assert np.array_equal(got.coverage, ref.coverage)
    assert np.array_equal(got.color, ref.color)
    if ref.depth is not None:
        assert np.array_equal(got.depth.view(np.uint32), ref.depth.view(np.uint32))
    fast = render_scene(dev, queue, sc, want_coverage=False, wgsl=text(name))
    assert np.array_equal(fast.color, ref.color)
