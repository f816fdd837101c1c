"""Execute WGSL snippets on the B200 through the emitter + NVRTC and compare with the known results
of the reference's own shader-compiler tests (naga-cranelift/src/tests.rs), see tests/wgsl_cases.py.
The value under test leaves the fragment stage through @builtin(frag_depth) into a Depth32Float
attachment (compare Always, write on), which keeps every bit of the f32."""
import numpy as np
import pytest

from tests.wgsl_cases import CASES, module_for

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from wgpu_cpu_b200 import api
    dev, queue = api.instance().request_adapter().request_device(0)
    return dev, queue


def evaluate(dev, queue, wgsl: str) -> float:
    module = dev.create_shader_module(wgsl)
    pipe = dev.create_render_pipeline(vertex_module=module, fragment_module=module,
                                      depth_stencil={"depth_compare": "always", "depth_write_enabled": True},
                                      targets=["rgba8unorm"])
    color = dev.create_texture(8, 8, "rgba8unorm")
    depth = dev.create_texture(8, 8, "depth32float")
    enc = dev.create_command_encoder()
    with enc.begin_render_pass([{"view": color.create_view(), "load": ("clear", (0, 0, 0, 0))}],
                               {"view": depth.create_view(), "depth_load": ("clear", 0.0)}) as rp:
        rp.set_pipeline(pipe)
        rp.draw(range(0, 3))
    idx = queue.submit([enc.finish()])
    dev.poll(True, idx)
    d = depth.read()
    c = color.read()
    assert (c[..., 0] == 255).all(), "the full-screen triangle must cover the target"
    assert (d == d[0, 0]).all()
    return float(d[0, 0])


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_wgsl_case(gpu, case):
    dev, queue = gpu
    got = evaluate(dev, queue, module_for(case))
    assert got == np.float32(case[4]), f"{case[0]}: got {got!r}, expected {case[4]!r}"
