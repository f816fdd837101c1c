"""examples/hello_mesh (C++ host over the C ABI, the reference's hello_mesh.rs flow) renders the teapot bit-exactly:
raw colour and depth against the CPU oracle, and the dumped PNG against the raw colour."""
import os
import subprocess

import numpy as np
import pytest

from wgpu_cpu_b200 import scenes as S

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_renders_hello_mesh(tmp_path):
    from oracle import pyoracle
    from tests.test_c_abi import _decode_png
    scene = S.hello_mesh(256, 192)
    ref = pyoracle.render(scene, want_coverage=False)
    (tmp_path / "v.bin").write_bytes(scene.vertex_buffers[0].tobytes())
    (tmp_path / "i.bin").write_bytes(scene.index_data.astype(np.uint32).tobytes())
    (tmp_path / "u.bin").write_bytes(np.ascontiguousarray(scene.bindings[(0, 0)][1]).tobytes())
    out = str(tmp_path / "frame")
    exe = os.path.join(ROOT, "examples", "hello_mesh")
    if os.environ.get("WGB_CUSIM") == "1":          # the same example source linked against the software-model build
        from tests.cusim import build as cusim_build
        exe = cusim_build.build_example()
    p = subprocess.run([exe, os.path.join(ROOT, "wgpu-cpu_b200", "shaders", "mesh_vertex_color.wgsl"),
                        str(tmp_path / "v.bin"), str(tmp_path / "i.bin"), str(tmp_path / "u.bin"), "256", "192", out],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert f"primitives {scene.num_primitives}" in p.stdout
    color = np.fromfile(out + ".rgba", dtype=np.uint8).reshape(192, 256, 4)
    depth = np.fromfile(out + ".depth", dtype=np.float32).reshape(192, 256)
    assert np.array_equal(color, ref.color)
    assert np.array_equal(depth.view(np.uint32), ref.depth.view(np.uint32))
    assert np.array_equal(_decode_png(out + ".png"), color)


def test_cpp_host_presents_hello_mesh_on_a_surface(tmp_path):
    """The windowed flow of hello_mesh.rs (:236-250, 380-410: create_surface, get_capabilities, formats[0], configure,
    get_current_texture -> pass -> submit -> present) against the headless surface: the window receives the oracle's
    Bgra8Unorm frame."""
    from oracle import pyoracle
    scene = S.hello_mesh(256, 192)
    scene.color_format = "bgra8unorm"
    ref = pyoracle.render(scene, want_coverage=False)
    (tmp_path / "v.bin").write_bytes(scene.vertex_buffers[0].tobytes())
    (tmp_path / "i.bin").write_bytes(scene.index_data.astype(np.uint32).tobytes())
    (tmp_path / "u.bin").write_bytes(np.ascontiguousarray(scene.bindings[(0, 0)][1]).tobytes())
    out = str(tmp_path / "frame")
    exe = os.path.join(ROOT, "examples", "hello_mesh")
    if os.environ.get("WGB_CUSIM") == "1":
        from tests.cusim import build as cusim_build
        exe = cusim_build.build_example()
    p = subprocess.run([exe, os.path.join(ROOT, "wgpu-cpu_b200", "shaders", "mesh_vertex_color.wgsl"),
                        str(tmp_path / "v.bin"), str(tmp_path / "i.bin"), str(tmp_path / "u.bin"), "256", "192", out, "3"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert "presented 3 frames of 256x192" in p.stdout
    window = np.fromfile(out + ".window", dtype=np.uint8).reshape(192, 256, 4)
    assert np.array_equal(window, ref.color)


def test_cpp_host_renders_hello_texture(tmp_path):
    """examples/hello_texture (the reference's hello_texture.rs flow: texture from image bytes, Repeat / Nearest sampler, a
    second bind group) renders the textured bunny bit-exactly."""
    from oracle import pyoracle
    from tests.test_c_abi import _decode_png
    scene = S.hello_texture(240, 160)
    ref = pyoracle.render(scene, want_coverage=False)
    image = np.ascontiguousarray(scene.bindings[(1, 0)][1])
    (tmp_path / "v.bin").write_bytes(scene.vertex_buffers[0].tobytes())
    (tmp_path / "i.bin").write_bytes(scene.index_data.astype(np.uint32).tobytes())
    (tmp_path / "u.bin").write_bytes(np.ascontiguousarray(scene.bindings[(0, 0)][1]).tobytes())
    (tmp_path / "t.rgba").write_bytes(image.tobytes())
    out = str(tmp_path / "frame")
    exe = os.path.join(ROOT, "examples", "hello_texture")
    if os.environ.get("WGB_CUSIM") == "1":
        from tests.cusim import build as cusim_build
        exe = cusim_build.build_example("hello_texture")
    p = subprocess.run([exe, os.path.join(ROOT, "wgpu-cpu_b200", "shaders", "mesh_textured.wgsl"), str(tmp_path / "v.bin"), str(tmp_path / "i.bin"),
                        str(tmp_path / "u.bin"), str(tmp_path / "t.rgba"), str(image.shape[1]), str(image.shape[0]), "240", "160", out],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert f"primitives {scene.num_primitives}" in p.stdout
    color = np.fromfile(out + ".rgba", dtype=np.uint8).reshape(160, 240, 4)
    depth = np.fromfile(out + ".depth", dtype=np.float32).reshape(160, 240)
    assert np.array_equal(color, ref.color)
    assert np.array_equal(depth.view(np.uint32), ref.depth.view(np.uint32))
    assert np.array_equal(_decode_png(out + ".png"), color)


def test_cpp_host_renders_hello_shader(tmp_path):
    """examples/hello_shader (hello_shader.rs: no vertex buffers, position and colour from vertex_index) renders the
    reference's test triangle bit-exactly, colour and depth; the depth dump is the 8-bit grey of lib.rs:129-158."""
    from oracle import pyoracle
    from tests.test_c_abi import _decode_png
    scene = S.colored_triangle("draw_backwards_no_cull", 200, 120)       # front face Ccw, no culling: hello_shader's pipeline state
    ref = pyoracle.render(scene, want_coverage=False)
    out = str(tmp_path / "frame")
    exe = os.path.join(ROOT, "examples", "hello_shader")
    if os.environ.get("WGB_CUSIM") == "1":
        from tests.cusim import build as cusim_build
        exe = cusim_build.build_example("hello_shader")
    p = subprocess.run([exe, os.path.join(ROOT, "wgpu-cpu_b200", "shaders", "index_triangle.wgsl"), "3", "200", "120", out],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    color = np.fromfile(out + ".rgba", dtype=np.uint8).reshape(120, 200, 4)
    depth = np.fromfile(out + ".depth", dtype=np.float32).reshape(120, 200)
    assert np.array_equal(color, ref.color)
    assert np.array_equal(depth.view(np.uint32), ref.depth.view(np.uint32))
    assert np.array_equal(_decode_png(out + ".png"), color)
    grey = np.clip(np.trunc(depth * np.float32(255.0)), 0, 255).astype(np.uint8)
    assert np.array_equal(_decode_png(out + ".depth.png")[:, :, 0], grey)


def test_cpp_host_presents_hello_shader_on_a_surface(tmp_path):
    """hello_shader's windowed flow (hello_shader.rs:168-185, 330-350) against the headless surface: a draw without vertex
    or index buffers into the surface's Bgra8Unorm texture reaches the window as the oracle renders it."""
    from oracle import pyoracle
    scene = S.colored_triangle("draw_backwards_no_cull", 200, 120)
    scene.color_format = "bgra8unorm"
    ref = pyoracle.render(scene, want_coverage=False)
    out = str(tmp_path / "frame")
    exe = os.path.join(ROOT, "examples", "hello_shader")
    if os.environ.get("WGB_CUSIM") == "1":
        from tests.cusim import build as cusim_build
        exe = cusim_build.build_example("hello_shader")
    p = subprocess.run([exe, os.path.join(ROOT, "wgpu-cpu_b200", "shaders", "index_triangle.wgsl"), "3", "200", "120", out, "2"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert "presented 2 frames" in p.stdout
    window = np.fromfile(out + ".window", dtype=np.uint8).reshape(120, 200, 4)
    assert np.array_equal(window, ref.color)


def test_cpp_host_presents_hello_texture_on_a_surface(tmp_path):
    """hello_texture's windowed flow against the headless surface: the textured bunny (sampled texture, Repeat / Nearest
    sampler, two bind groups) through the surface's Bgra8Unorm texture, byte-exact in the window."""
    from oracle import pyoracle
    scene = S.hello_texture(240, 160)
    scene.color_format = "bgra8unorm"
    ref = pyoracle.render(scene, want_coverage=False)
    image = np.ascontiguousarray(scene.bindings[(1, 0)][1])
    (tmp_path / "v.bin").write_bytes(scene.vertex_buffers[0].tobytes())
    (tmp_path / "i.bin").write_bytes(scene.index_data.astype(np.uint32).tobytes())
    (tmp_path / "u.bin").write_bytes(np.ascontiguousarray(scene.bindings[(0, 0)][1]).tobytes())
    (tmp_path / "t.rgba").write_bytes(image.tobytes())
    out = str(tmp_path / "frame")
    exe = os.path.join(ROOT, "examples", "hello_texture")
    if os.environ.get("WGB_CUSIM") == "1":
        from tests.cusim import build as cusim_build
        exe = cusim_build.build_example("hello_texture")
    p = subprocess.run([exe, os.path.join(ROOT, "wgpu-cpu_b200", "shaders", "mesh_textured.wgsl"), str(tmp_path / "v.bin"), str(tmp_path / "i.bin"),
                        str(tmp_path / "u.bin"), str(tmp_path / "t.rgba"), str(image.shape[1]), str(image.shape[0]), "240", "160", out, "2"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert "presented 2 frames" in p.stdout
    window = np.fromfile(out + ".window", dtype=np.uint8).reshape(160, 240, 4)
    assert np.array_equal(window, ref.color)
