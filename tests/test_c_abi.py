"""The C-ABI library loads, exports every symbol include/wgpu_b200.h declares, and its host-side logic
(object lifetime, validation, error reporting, recording) behaves -- all without a GPU, using the
compile-only device (translation + NVRTC, no execution).  No compute call is made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from wgpu_cpu_b200 import api, shaders

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "wgpu_b200.h")).read()
    return sorted(set(re.findall(r"WGB_API\s+[\w\s\*]+?\b(wgb_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = api.load_library()
    syms = declared_symbols()
    assert len(syms) >= 50
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"missing exports: {missing}"
    assert lib.wgb_version().decode().endswith("sm_100a")


def test_no_device_means_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    adapter = api.instance().request_adapter()
    assert adapter.get_info()["cuda_device_count"] == 0
    with pytest.raises(api.WgpuError) as e:
        adapter.request_device()
    assert "no CPU fallback" in str(e.value)


@pytest.fixture(scope="module")
def offline():
    dev, queue = api.instance().request_adapter().request_device(api.CUDA_DEVICE_COMPILE_ONLY)
    return dev, queue


def test_buffer_map_write_read_roundtrip(offline):
    dev, _ = offline
    b = dev.create_buffer(64, api.BUFFER_USAGE["VERTEX"], mapped_at_creation=True)
    view = b.get_mapped_range()
    assert view.shape == (64,) and not view.any()          # zero-initialised like Vec<u8> (buffer.rs:31-35)
    view[:] = np.arange(64, dtype=np.uint8)
    assert b.get_mapped_range(16, 8).tolist() == list(range(16, 24))
    b.unmap()
    with pytest.raises(api.WgpuError):
        b.get_mapped_range()                               # not mapped any more
    with pytest.raises(api.WgpuError):
        dev.create_buffer(16, mapped_at_creation=True).get_mapped_range(8, 16)   # out of range


def test_pipeline_creation_compiles_wgsl_with_nvrtc(offline):
    dev, _ = offline
    m = dev.create_shader_module(shaders.wgsl("hello_mesh"))
    p = dev.create_render_pipeline(
        vertex_module=m, fragment_module=m, front_face="cw", cull_mode="back", depth_stencil={"depth_compare": "less"},
        vertex_buffers=[{"array_stride": 32, "attributes": [("float32x4", 0, 0), ("float32x4", 16, 1)]}], targets=["rgba8unorm-srgb"])
    src = p.get_source()
    assert "#define WGB_RESOLVE 1" in src and "#define WGB_CULL 2" in src and "#define WGB_FRONT_FACE_CW 1" in src
    assert "#define WGB_ATTR1_OFFSET 16u" in src and '#include "wgb_raster.cuh"' in src


def test_shader_errors_surface_at_pipeline_creation(offline):
    dev, _ = offline
    m = dev.create_shader_module("@vertex fn vs_main() -> @builtin(position) vec4f { return vec4f(q); }")
    with pytest.raises(api.WgpuError) as e:
        dev.create_render_pipeline(vertex_module=m, targets=[])
    assert "unknown identifier 'q'" in str(e.value)
    # a missing vertex attribute is a compile error of the generated translation unit
    m = dev.create_shader_module(shaders.wgsl("hello_mesh"))
    with pytest.raises(api.WgpuError) as e:
        dev.create_render_pipeline(vertex_module=m, fragment_module=m, targets=["rgba8unorm"])
    assert e.value.status == 5 and "WGB_ATTR" in str(e.value)


def test_unsupported_state_is_reported(offline):
    dev, _ = offline
    m = dev.create_shader_module(shaders.wgsl("colored_triangle"))
    with pytest.raises(api.WgpuError) as e:      # state.rs:438-478 panics for PolygonMode::Line
        dev.create_render_pipeline(vertex_module=m, fragment_module=m, targets=["rgba8unorm"], polygon_mode=1)
    assert e.value.status == 2
    # NotEqual + depth write has no closed form: it takes the ordered tile kernel, which runs every topology
    for topology in ("triangle-list", "line-list", "point-list"):
        p = dev.create_render_pipeline(vertex_module=m, fragment_module=m, targets=["rgba8unorm"], topology=topology,
                                       depth_stencil={"depth_compare": "not-equal", "depth_write_enabled": True})
        assert "#define WGB_RESOLVE 7" in p.get_source()
    with pytest.raises(api.WgpuError):           # binding.rs:161 todo!()
        dev.create_sampler(address_mode_u="clamp-to-border")
    with pytest.raises(api.WgpuError):
        dev.create_texture(0, 4, "rgba8unorm")


def test_recording_and_submit_on_compile_only_device(offline):
    dev, queue = offline
    m = dev.create_shader_module(shaders.wgsl("colored_triangle"))
    p = dev.create_render_pipeline(vertex_module=m, fragment_module=m, targets=["rgba8unorm"])
    tex = dev.create_texture(8, 8, "rgba8unorm")
    enc = dev.create_command_encoder()
    rp = enc.begin_render_pass([{"view": tex.create_view(), "load": ("clear", (0, 0, 0, 1))}])
    rp.set_pipeline(p)
    rp.set_viewport(0, 0, 8, 8)
    rp.set_scissor_rect(0, 0, 8, 8)
    rp.set_blend_constant((0, 0, 0, 0))
    rp.set_stencil_reference(1)
    rp.draw(range(0, 3))
    rp.end()
    rp.end()                                     # end() is idempotent (also called from Drop, mod.rs:325-329)
    with pytest.raises(api.WgpuError):
        rp.draw(range(0, 3))                     # recording after end
    cb = enc.finish()
    with pytest.raises(api.WgpuError):
        enc.finish()
    with pytest.raises(api.WgpuError) as e:      # there is no CPU path to execute it on
        queue.submit([cb])
    assert "compile-only" in str(e.value)
    assert dev.poll(True) == 1                   # QueueEmpty (device.rs:266-268)


def test_band_rows_match_python(offline):
    from wgpu_cpu_b200.multigpu import band_rows
    dev, _ = offline
    for h in (1, 31, 32, 33, 2160, 4320, 1080):
        for n in (1, 2, 3, 4, 8):
            covered = []
            for r in range(n):
                dev.set_band(r, n)
                assert dev.band_rows(h) == band_rows(h, r, n)
                covered.append(dev.band_rows(h))
            assert covered[0][0] == 0 and covered[-1][1] == h
            assert all(covered[i][1] == covered[i + 1][0] for i in range(n - 1))
    dev.set_band(0, 1)
    with pytest.raises(api.WgpuError):
        dev.set_band(2, 2)


def test_encoder_copy_validation(offline):
    dev, _ = offline
    buf = dev.create_buffer(1024)
    tex = dev.create_texture(16, 16, "rgba8unorm")
    enc = dev.create_command_encoder()
    enc.copy_texture_to_buffer(tex, buf)                       # 16*16*4 = 1024 bytes: fits exactly
    enc.copy_buffer_to_buffer(buf, 0, buf, 512, 512)
    enc.clear_buffer(buf, 16, 32)
    with pytest.raises(api.WgpuError):
        enc.copy_texture_to_buffer(tex, buf, bytes_per_row=256)  # 15*256 + 64 > 1024
    with pytest.raises(api.WgpuError):
        enc.copy_texture_to_buffer(tex, buf, size=(17, 1))
    with pytest.raises(api.WgpuError):
        enc.copy_buffer_to_buffer(buf, 1000, buf, 0, 100)
    with pytest.raises(api.WgpuError):
        enc.copy_texture_to_buffer(tex, buf, bytes_per_row=32)   # pitch smaller than a row
    enc.finish()
    with pytest.raises(api.WgpuError):
        enc.clear_buffer(buf)                                    # encoder already finished


def _decode_png(path):
    """Minimal PNG reader for the files the library writes (8-bit, filter 0 rows)."""
    import struct
    import zlib
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, hdr = 8, b"", None
    while pos < len(raw):
        n, typ = struct.unpack(">I4s", raw[pos:pos + 8])
        body = raw[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(typ + body)
        if typ == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"IDAT":
            idat += body
        pos += 12 + n
    w, h, depth, ctype = hdr[:4]
    assert depth == 8
    ch = {0: 1, 2: 3, 6: 4}[ctype]
    rows = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + w * ch)
    assert not rows[:, 0].any()
    return rows[:, 1:].reshape(h, w, ch)


@pytest.mark.parametrize("shape", [(5, 7, 4), (300, 301, 3), (33, 2)])
def test_png_writer_roundtrip(tmp_path, shape):
    """wgb_write_png (behind wgb_texture_dump_png, the reference's dump_texture lib.rs:111-158): stored-deflate PNG,
    valid CRCs and Adler-32, more than one 64 KiB block."""
    from wgpu_cpu_b200 import api
    px = np.random.default_rng(1).integers(0, 256, shape, dtype=np.uint8)
    path = str(tmp_path / "t.png")
    api.write_png(path, px)
    got = _decode_png(path)
    assert np.array_equal(got.reshape(px.shape), px)


def test_unknown_feature_bits_are_rejected():
    adapter = api.instance().request_adapter()
    with pytest.raises(api.WgpuError):
        adapter.request_device(api.CUDA_DEVICE_COMPILE_ONLY, features=1 << 20)
    dev, _ = adapter.request_device(api.CUDA_DEVICE_COMPILE_ONLY, features=sum(api.FEATURE.values()))
    assert dev is not None


def test_ordered_variants_compile_offline():
    """The ordered tile kernel (NotEqual + depth write in parity mode; blending on a WGB_FEATURE_BLEND device) is a
    separate variant of the pipeline translation unit: translate + NVRTC-compile it without a GPU."""
    adapter = api.instance().request_adapter()
    dev, _ = adapter.request_device(api.CUDA_DEVICE_COMPILE_ONLY, features=api.FEATURE["BLEND"] | api.FEATURE["COLOR_WRITE_MASK"])
    m = dev.create_shader_module(shaders.wgsl("hello_mesh"))
    vbs = [{"array_stride": 32, "attributes": [("float32x4", 0, 0), ("float32x4", 16, 1)]}]
    blend = {"color": ("src-alpha", "one-minus-src-alpha", "add"), "alpha": ("one", "one-minus-dst-alpha", "reverse-subtract")}
    p = dev.create_render_pipeline(vertex_module=m, fragment_module=m, vertex_buffers=vbs, depth_stencil={"depth_compare": "less"},
                                   targets=[{"format": "bgra8unorm-srgb", "blend": blend, "write_mask": 7}])
    src = p.get_source()
    assert "#define WGB_RESOLVE 7" in src
    assert "t == 0 ? (f == 0 ? 1 : f == 1 ? 4 : f == 2 ? 5 : f == 3 ? 0 : f == 4 ? 1 : f == 5 ? 9 : 2)" in src
    # the same state on a device without the feature: blend ignored, closed-form kernel (Less + write = MIN_FIRST)
    dev0, _ = adapter.request_device(api.CUDA_DEVICE_COMPILE_ONLY)
    m0 = dev0.create_shader_module(shaders.wgsl("hello_mesh"))
    p0 = dev0.create_render_pipeline(vertex_module=m0, fragment_module=m0, vertex_buffers=vbs, depth_stencil={"depth_compare": "less"},
                                     targets=[{"format": "bgra8unorm-srgb", "blend": blend, "write_mask": 7}])
    assert "#define WGB_RESOLVE 1" in p0.get_source()
    pn = dev0.create_render_pipeline(vertex_module=m0, fragment_module=m0, vertex_buffers=vbs, targets=["rgba8unorm"],
                                     depth_stencil={"depth_compare": "not-equal", "depth_write_enabled": True})
    assert "#define WGB_RESOLVE 7" in pn.get_source()
    # an early depth test ahead of discard / frag_depth, or followed by the late test (Allow), has no closed form either;
    # compiled here for sm_100a for a triangle, a line and a point topology
    for name, topology in (("early_force", "triangle-list"), ("early_allow", "line-strip"), ("early_force", "point-list")):
        me = dev0.create_shader_module(shaders.wgsl(name))
        pe = dev0.create_render_pipeline(vertex_module=me, fragment_module=me, vertex_buffers=vbs, targets=["rgba8unorm"], topology=topology,
                                         depth_stencil={"depth_compare": "less", "depth_write_enabled": True})
        assert "#define WGB_RESOLVE 7" in pe.get_source() and "#define WGB_FS_EARLY_DEPTH" in pe.get_source()
    # without a depth test the attribute changes nothing: closed-form kernel
    me = dev0.create_shader_module(shaders.wgsl("early_force"))
    assert "#define WGB_RESOLVE 0" in dev0.create_render_pipeline(vertex_module=me, fragment_module=me, vertex_buffers=vbs, targets=["rgba8unorm"]).get_source()


def test_multiple_colour_targets_compile_offline(offline):
    """Three colour attachments: the closed-form and the ordered tile kernel compile for sm_100a with WGB_NUM_COLOR 3."""
    dev, _ = offline
    m = dev.create_shader_module(shaders.wgsl("mrt"))
    vbs = [{"array_stride": 32, "attributes": [("float32x4", 0, 0), ("float32x4", 16, 1)]}]
    for compare, resolve in (("less", 1), ("not-equal", 7)):
        p = dev.create_render_pipeline(vertex_module=m, fragment_module=m, vertex_buffers=vbs, targets=["rgba8unorm", "bgra8unorm", "rg8unorm"],
                                       depth_stencil={"depth_compare": compare, "depth_write_enabled": True})
        src = p.get_source()
        assert "#define WGB_NUM_COLOR 3" in src and f"#define WGB_RESOLVE {resolve}" in src and "#define WGB_FS_COLOR_MASK 7" in src


def test_header_is_plain_c_and_a_c_host_links(tmp_path):
    """The boundary is a C ABI: include/wgpu_b200.h compiles as C99 (-pedantic), and a host written in C links against
    the library and walks instance -> adapter -> surface without a GPU (no compute call)."""
    import subprocess
    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "wgpu_b200.h"
int main(void) {
    wgb_instance inst = NULL; wgb_adapter adapter = NULL; wgb_surface surface = NULL;
    wgb_surface_capabilities caps; wgb_adapter_info info; int32_t ok = 0; wgb_texture t = NULL; uint32_t st = 0;
    if (wgb_create_instance(NULL, &inst) != WGB_OK) return 1;
    if (wgb_instance_request_adapter(inst, &adapter) != WGB_OK) return 2;
    if (wgb_adapter_get_info(adapter, &info) != WGB_OK) return 3;
    if (wgb_instance_create_surface(inst, NULL, &surface) != WGB_OK) return 4;
    if (wgb_adapter_is_surface_supported(adapter, surface, &ok) != WGB_OK || !ok) return 5;
    if (wgb_surface_get_capabilities(surface, adapter, &caps) != WGB_OK) return 6;
    if (caps.format_count != 1 || caps.formats[0] != WGB_TEXTURE_FORMAT_BGRA8_UNORM || caps.present_modes[0] != WGB_PRESENT_MODE_IMMEDIATE) return 7;
    if (wgb_surface_get_current_texture(surface, &t, &st) != WGB_ERROR_VALIDATION) return 8;       /* not configured yet */
    if (!strstr(wgb_last_error(), "Surface not configured yet")) return 9;
    printf("%s\n", wgb_version());
    wgb_release((wgb_object)surface); wgb_release((wgb_object)adapter); wgb_release((wgb_object)inst);
    return 0;
}
''')
    exe = tmp_path / "host"
    lib_dir = os.path.dirname(api.LIB_PATH)
    lib = os.path.basename(api.LIB_PATH)[3:-3]
    p = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{os.path.join(ROOT, 'include')}", str(src), "-o", str(exe),
                        f"-L{lib_dir}", f"-l{lib}", f"-Wl,-rpath,{lib_dir}"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, (r.returncode, r.stderr)
    assert "sm_100a" in r.stdout
