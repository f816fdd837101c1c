"""WGSL -> CUDA C++ emitter, host side (no GPU): golden emitter output for the built-in shaders,
translation of every snippet case, and the errors raised for WGSL the backend does not run."""
import pytest

from tests.wgsl_cases import CASES, module_for
from wgpu_cpu_b200 import api, shaders


@pytest.mark.parametrize("name", shaders.NAMES)
@pytest.mark.parametrize("stage,entry,tag", [(api.STAGE_VERTEX, "vs_main", "vs"), (api.STAGE_FRAGMENT, "fs_main", "fs")])
def test_golden_emitter_output(name, stage, entry, tag):
    assert api.translate_wgsl(shaders.wgsl(name), stage, entry) == shaders.emitted(name, tag)


def test_arithmetic_contract_in_emitted_text():
    """One IEEE operation per operator, never a*b+c: scalar float operators become wgb_* calls, and the
    camera transform is the prelude's column-accumulating operator* (binary.rs:297-323)."""
    vs = shaders.emitted("hello_mesh", "vs")
    assert "wgb_load<mat4x4f>(wgb, 0, 0, 0u) * object_position" in vs
    fs = shaders.emitted("procedural", "fs")
    assert "wgb_add(wgb_sub(wgb_mul(zx, zx), wgb_mul(zy, zy)), cx)" in fs
    assert "wgb_add(wgb_mul(wgb_mul(2.0f, zx), zy), cy)" in fs


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_snippets_translate(case):
    src = module_for(case)
    assert "wgb_vs_entry" in api.translate_wgsl(src, api.STAGE_VERTEX, "vs_main")
    fs = api.translate_wgsl(src, api.STAGE_FRAGMENT, "fs_main")
    assert "#define WGB_FS_WRITES_FRAG_DEPTH 1" in fs and "#define WGB_FS_COLOR_MASK 1" in fs


def _fs(body, decls="", ret="@location(0) vec4f"):
    return f"{decls}\n@fragment fn fs_main(@builtin(position) p: vec4f) -> {ret} {{ {body} }}"


def test_inter_stage_layout_and_interpolation():
    src = """
struct VOut { @builtin(position) p: vec4f, @location(2) @interpolate(flat) id: u32, @location(0) uv: vec2f,
              @location(1) @interpolate(linear) n: vec3f, }
@vertex fn vs_main(@location(0) a: vec4f) -> VOut { return VOut(a, 1u, vec2f(0.0), vec3f(1.0)); }
@fragment fn fs_main(v: VOut) -> @location(0) vec4f { return vec4f(v.uv, f32(v.id), v.n.x); }
"""
    vs = api.translate_wgsl(src, api.STAGE_VERTEX, "vs_main")
    # locations packed in declaration order with naga's alignments (bindings.rs:307-346): u32 @0, vec2 @2 (8-byte), vec3 @4 (16-byte)
    assert "#define WGB_VS_LOC2_SLOT 0" in vs and "#define WGB_VS_LOC0_SLOT 2" in vs and "#define WGB_VS_LOC1_SLOT 4" in vs
    assert "#define WGB_VS_VARYING_SLOTS 7" in vs
    fs = api.translate_wgsl(src, api.STAGE_FRAGMENT, "fs_main")
    assert "WGB_VS_LOC2_SLOT + 1) ? 0" in fs      # flat
    assert "WGB_VS_LOC0_SLOT + 2) ? 2" in fs      # default = perspective
    assert "WGB_VS_LOC1_SLOT + 3) ? 1" in fs      # linear


def test_frag_depth_after_location_is_ignored():
    """fragment.rs:457-488: the late depth test runs at the first @location output, so a frag_depth
    declared after it never reaches the test."""
    src = _fs("return O(vec4f(1.0), 0.25);", "struct O { @location(0) c: vec4f, @builtin(frag_depth) d: f32, }", "O")
    assert "#define WGB_FS_WRITES_FRAG_DEPTH 0" in api.translate_wgsl(src, api.STAGE_FRAGMENT, "fs_main")


def test_discard_propagates_through_calls():
    src = _fs("maybe(p.x); return vec4f(1.0);", "fn maybe(x: f32) { if (x > 3.0) { discard; } }")
    fs = api.translate_wgsl(src, api.STAGE_FRAGMENT, "fs_main")
    assert "#define WGB_FS_MAY_DISCARD 1" in fs and "if (wgb_inv.killed) return vec4f();" in fs


def test_uniform_layout_offsets():
    src = """
struct Light { dir: vec3f, power: f32, color: vec3f, }
struct U { m: mat4x4f, lights: array<Light, 2>, k: vec2f, }
@group(0) @binding(1) var<uniform> u: U;
@fragment fn fs_main(@builtin(position) p: vec4f) -> @location(0) vec4f {
    let i = u32(p.x);
    return vec4f(u.lights[1].color, u.lights[i].power + u.k.y);
}"""
    fs = api.translate_wgsl(src, api.STAGE_FRAGMENT, "fs_main")
    assert "wgb_load<vec3f>(wgb, 0, 1, 112u)" in fs          # 64 + 32*1 + 16
    assert "wgb_load<f32>(wgb, 0, 1, 76u + min((u32)(i), 1u) * 32u)" in fs
    assert "wgb_load<f32>(wgb, 0, 1, 132u)" in fs            # k at 128, .y


@pytest.mark.parametrize("src,fragment", [
    (_fs("return vec4f(1.0);", "override k: f32 = 1.0;"), "override"),
    ("@compute @workgroup_size(1) fn fs_main() {}", "not a fragment entry point"),
    (_fs("return textureSample(t, s, p.xy, vec2i(1, 1));", "@group(0) @binding(0) var t: texture_2d<f32>;\n@group(0) @binding(1) var s: sampler;"), "offset"),
    (_fs("b = 1.0; return vec4f(1.0);", "@group(0) @binding(0) var<storage, read_write> b: f32;"), "read-only"),
    (_fs("return vec4f(q);"), "unknown identifier"),
    (_fs("let a: u32 = 1.5; return vec4f(1.0);"), "cannot convert"),
    (_fs("return vec4f(1.0) * mat4x4f();"), "vector * matrix"),
    (_fs("return vec4f(1.0);", "var<workgroup> w: f32;"), "workgroup"),
    (_fs("return vec4f(1.0);", "", "@location(0) vec3f"), "vec4<f32>"),
])
def test_unsupported_wgsl_is_rejected_with_a_message(src, fragment):
    with pytest.raises(api.WgpuError) as e:
        api.translate_wgsl(src, api.STAGE_FRAGMENT, "fs_main")
    assert fragment in str(e.value)


def test_f16_values_translate_and_f16_at_the_boundaries_is_unsupported():
    """f16 (naga-cranelift/src/types.rs:103-136 has the scalar; none of the reference's tests or examples uses it): values,
    vectors, literals, conversions, arithmetic, comparisons, select, abs / min / max / clamp translate -- every operator
    one correctly rounded binary16 operation (wgb_prelude.cuh).  f16 in buffers, in entry-point interfaces, in matrices
    and in the other math builtins is WGB_ERROR_UNSUPPORTED (2); a malformed shader stays WGB_ERROR_SHADER (5)."""
    fs = api.translate_wgsl("enable f16;\n" + _fs("var a: f16 = 1.5h; let v = vec2h(a, 2.0h) * 3.0h; let c = clamp(-v.x, 0.0h, v.y); return vec4f(f32(c), f32(v.y % a), f32(a < c), 1.0);"),
                            api.STAGE_FRAGMENT, "fs_main")
    assert "f16 a = wgb_f16(1.5f);" in fs and "vec2h" in fs and "wgb_clamp(" in fs and "wgb_rem(" in fs
    for decl, body in (("@group(0) @binding(0) var<uniform> u: f16;", "return vec4f(f32(u));"),
                       ("", "return vec4f(f32(sqrt(2.0h)));"),
                       ("", "let m = mat2x2<f16>(); return vec4f(1.0);")):
        with pytest.raises(api.WgpuError) as e:
            api.translate_wgsl("enable f16;\n" + _fs(body, decl), api.STAGE_FRAGMENT, "fs_main")
        assert e.value.status == 2 and "f16" in str(e.value), str(e.value)
    with pytest.raises(api.WgpuError) as e:
        api.translate_wgsl("enable f16;\n" + _fs("return vec4h(1.0h);", "", "@location(0) vec4<f16>"), api.STAGE_FRAGMENT, "fs_main")
    assert e.value.status == 2
    with pytest.raises(api.WgpuError) as e:
        api.translate_wgsl(_fs("return vec4f(q);"), api.STAGE_FRAGMENT, "fs_main")
    assert e.value.status == 5


def test_array_length_translates():
    """Expression::ArrayLength (tests.rs:997-1103): (bound size - array offset) / element stride."""
    from wgpu_cpu_b200 import api
    src = """
struct Output { @builtin(position) p: vec4f, @location(0) @interpolate(flat) output: u32, }
@group(0) @binding(0) var<storage, read> what_len_is_this_array: array<i32>;
struct Tail { head: vec4f, items: array<vec3f>, }
@group(1) @binding(2) var<storage, read> tail: Tail;
@vertex fn main() -> Output { return Output(vec4f(), arrayLength(&what_len_is_this_array) + arrayLength(&tail.items)); }
"""
    cu = api.translate_wgsl(src, api.STAGE_VERTEX, "main")
    assert "wgb_array_length(wgb, 0, 0, 0u, 4u)" in cu and "wgb_array_length(wgb, 1, 2, 16u, 16u)" in cu
    assert "struct Tail { vec4f head; };" in cu
    import pytest
    with pytest.raises(api.WgpuError):
        api.translate_wgsl(src.replace("arrayLength(&tail.items)", "arrayLength(&tail.head)"), api.STAGE_VERTEX, "main")


def test_only_the_entry_points_call_graph_is_emitted():
    """A module shared by both stages may hold helpers that are valid in one stage only: a helper that discards must not
    stop the vertex stage of the same module from translating (it is outside that entry point's call graph), while a
    vertex entry point that does call it is an error."""
    src = """
fn cut(x: f32) -> f32 { if (x > 3.0) { discard; } return x; }
fn unused_everywhere() -> f32 { return 1.0; }
fn shared_helper(x: f32) -> f32 { return x * 2.0; }
@vertex fn vs_main(@location(0) a: vec4f) -> @builtin(position) vec4f { return vec4f(shared_helper(a.x), a.y, a.z, a.w); }
@fragment fn fs_main(@builtin(position) p: vec4f) -> @location(0) vec4f { return vec4f(cut(shared_helper(p.x))); }
"""
    vs = api.translate_wgsl(src, api.STAGE_VERTEX, "vs_main")
    assert "shared_helper(" in vs and "cut(" not in vs and "unused_everywhere" not in vs
    fs = api.translate_wgsl(src, api.STAGE_FRAGMENT, "fs_main")
    assert "cut(" in fs and "shared_helper(" in fs and "unused_everywhere" not in fs and "#define WGB_FS_MAY_DISCARD 1" in fs
    bad = src.replace("vec4f(shared_helper(a.x), a.y", "vec4f(cut(a.x), a.y")
    with pytest.raises(api.WgpuError) as e:
        api.translate_wgsl(bad, api.STAGE_VERTEX, "vs_main")
    assert "discard" in str(e.value)
