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


def evaluate(dev, queue, wgsl: str, bind_group_entries=None) -> float:
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
        if bind_group_entries:
            rp.set_bind_group(0, dev.create_bind_group(None, bind_group_entries))
        rp.draw(range(0, 3))
    idx = queue.submit([enc.finish()])
    dev.poll(True, idx)
    d = depth.read()
    c = color.read()
    assert (c[..., 0] == 255).all(), "the full-screen triangle must cover the target"
    assert (d == d[0, 0]).all()
    return float(d[0, 0])


def evaluate_batch(dev, queue, cases) -> dict:
    """All cases of a group in one module and one draw: column k of a (len x 4) target holds case k's value."""
    from tests.wgsl_cases import module_for_batch
    module = dev.create_shader_module(module_for_batch(cases))
    pipe = dev.create_render_pipeline(vertex_module=module, fragment_module=module,
                                      depth_stencil={"depth_compare": "always", "depth_write_enabled": True},
                                      targets=["rgba8unorm"])
    w, h = len(cases), 4
    color = dev.create_texture(w, h, "rgba8unorm")
    depth = dev.create_texture(w, h, "depth32float")
    enc = dev.create_command_encoder()
    with enc.begin_render_pass([{"view": color.create_view(), "load": ("clear", (0, 0, 0, 0))}],
                               {"view": depth.create_view(), "depth_load": ("clear", 0.0)}) as rp:
        rp.set_pipeline(pipe)
        rp.draw(range(0, 6))
    dev.poll(True, queue.submit([enc.finish()]))
    d = depth.read()
    assert (color.read()[..., 0] == 255).all(), "the two triangles must cover the target"
    assert (d == d[0:1, :]).all(), "every row evaluates the same cases"
    return {c[0]: float(d[0, k]) for k, c in enumerate(cases)}


@pytest.fixture(scope="module")
def case_values(gpu):
    """name -> value for every case, one shader module per group of non-colliding cases (4 NVRTC compilations instead of
    one per case); a group that fails as a whole is re-run case by case so that the failure names its case."""
    from tests.wgsl_cases import batches
    dev, queue = gpu
    values = {}
    for group in batches():
        try:
            values.update(evaluate_batch(dev, queue, group))
        except Exception:      # noqa: BLE001 -- fall back to one module per case
            for case in group:
                try:
                    values[case[0]] = evaluate(dev, queue, module_for(case))
                except Exception as e:      # noqa: BLE001
                    values[case[0]] = e
    return values


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_wgsl_case(case_values, case):
    got = case_values[case[0]]
    if isinstance(got, Exception):
        raise got
    assert got == np.float32(case[4]), f"{case[0]}: got {got!r}, expected {case[4]!r}"


def test_a_case_on_its_own_module(gpu):
    """The one-module-per-case harness the batches replace (kept for the fall-back path): same value."""
    dev, queue = gpu
    case = next(c for c in CASES if c[0] == "for_factorial")
    assert evaluate(dev, queue, module_for(case)) == np.float32(case[4])


def test_array_length_of_storage_bindings(gpu):
    """tests.rs:997-1103 `array_length`: a storage buffer of 123 i32 reports 123; also the runtime-sized tail of a
    struct (offset 16, vec3f stride 16) and element reads through both."""
    from wgpu_cpu_b200 import api
    dev, queue = gpu
    ints = np.full(123, 42, dtype=np.int32)
    ints[2] = 7
    tail = np.zeros(4 + 4 * 9, dtype=np.float32)            # head vec4f + 9 x vec3f (stride 16)
    tail[4 + 4 * 1 + 1] = 5.0                               # items[1].y
    a = dev.create_buffer_init(ints, api.BUFFER_USAGE["STORAGE"])
    b = dev.create_buffer_init(tail, api.BUFFER_USAGE["STORAGE"])
    case = ("array_length",
            "@group(0) @binding(0) var<storage, read> what_len_is_this_array: array<i32>;\n"
            "struct Tail { head: vec4f, items: array<vec3f>, }\n"
            "@group(0) @binding(1) var<storage, read> tail: Tail;",
            "let n = arrayLength(&what_len_is_this_array); let m = arrayLength(&tail.items);",
            "f32(n) + 1000.0 * f32(m) + 0.25 * tail.items[1].y + 0.125 * f32(what_len_is_this_array[2])", 0.0)
    got = evaluate(dev, queue, module_for(case), [{"binding": 0, "buffer": a}, {"binding": 1, "buffer": b}])
    assert got == np.float32(123 + 9000 + 1.25 + 0.875)


def test_texture_queries_loads_and_whole_value_uniform_loads(gpu):
    """textureDimensions / textureLoad (image.rs:107,119 are todo!() in the reference) and loads of a whole struct and a
    whole array out of a uniform buffer (each member / element fetched at its layout offset)."""
    from wgpu_cpu_b200 import api
    dev, queue = gpu
    img = np.zeros((3, 5, 4), dtype=np.uint8)               # 5 wide, 3 high
    img[2, 4] = [255, 51, 102, 255]
    tex = dev.create_texture_with_data(queue, 5, 3, "rgba8unorm", img)
    data = np.zeros(4 + 4 * 2 + 4 * 3, dtype=np.float32)    # m: vec4f | lights: 2 x {dir vec3f, power f32} | table: array<vec4f, 3>
    data[0:4] = [1.0, 2.0, 3.0, 4.0]
    data[4:8] = [0.5, 0.25, 0.125, 7.0]
    data[8:12] = [9.0, 8.0, 7.0, 11.0]
    data[12:24] = np.arange(12, dtype=np.float32)
    ub = dev.create_buffer_init(data, api.BUFFER_USAGE["UNIFORM"])
    case = ("texture_and_whole_loads",
            "struct Light { dir: vec3f, power: f32, }\nstruct U { m: vec4f, lights: array<Light, 2>, table: array<vec4f, 3>, }\n"
            "@group(0) @binding(0) var<uniform> u: U;\n@group(0) @binding(1) var t: texture_2d<f32>;",
            "let d = textureDimensions(t); let px = textureLoad(t, vec2i(4, 2), 0); let light = u.lights[1]; let table = u.table; var k = 2;",
            "f32(d.x) * 1000.0 + f32(d.y) * 100.0 + px.x + px.y * 5.0 + light.dir.y + light.power + table[k].w + table[0].y", 0.0)
    got = evaluate(dev, queue, module_for(case), [{"binding": 0, "buffer": ub}, {"binding": 1, "texture_view": tex.create_view()}])
    want = np.float32(5000.0 + 300.0)
    for term in (1.0, np.float32(np.float32(51.0 / 255.0) * np.float32(5.0)), 8.0, 11.0, 11.0, 1.0):
        want = np.float32(want + np.float32(term))
    assert got == want, (got, want)
