"""rust/wgpu-b200-sys (the FFI crate of INTEGRATION.md) is generated from include/wgpu_b200.h: it must be up to date
and declare every entry point and every descriptor field.  (There is no Rust toolchain in this image, so the crate is
not compiled here; the same header is what the C++ and Python hosts are tested against.)"""
import importlib.util
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gen():
    spec = importlib.util.spec_from_file_location("gen_rust_sys", os.path.join(ROOT, "tools", "gen_rust_sys.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_sys_crate_is_up_to_date_and_complete():
    gen = _gen()
    text = gen.generate()
    assert text == open(os.path.join(ROOT, "rust", "wgpu-b200-sys", "src", "lib.rs")).read(), "run python tools/gen_rust_sys.py"
    header = gen.strip_comments(open(os.path.join(ROOT, "include", "wgpu_b200.h")).read())
    declared = set(re.findall(r"WGB_API\s+[\w\s\*]+?\b(wgb_\w+)\s*\(", header))
    bound = set(re.findall(r"pub fn (wgb_\w+)\(", text))
    assert declared == bound and len(bound) >= 60
    # descriptor structs keep their field order and count
    for body, name in re.findall(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", header, flags=re.S):
        n_c = sum(len(d.split(",")) for d in body.split(";") if d.strip())
        rust = re.search(r"pub struct %s \{(.*?)\n\}" % name, text, flags=re.S).group(1)
        assert rust.count("pub ") == n_c, name
    assert "pub device_type: u32" in text and "pub name: [c_char; 128]" in text
    assert "pub fn wgb_queue_submit(queue: wgb_queue, command_buffers: *const wgb_command_buffer, count: u32, out_submission_index: *mut u64) -> wgb_status;" in text


def _split_top_level(text):
    parts, depth, cur = [], 0, ""
    for ch in text:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur)
    return parts


def _balanced(text, start, open_ch, close_ch):
    depth, i = 0, start
    while True:
        if text[i] == open_ch:
            depth += 1
        elif text[i] == close_ch:
            depth -= 1
            if depth == 0:
                return text[start + 1:i]
        i += 1


def test_backend_crate_matches_the_sys_crate():
    """rust/wgpu-b200 (the wgpu custom-backend traits as thin callers of the C ABI) cannot be compiled here; this keeps
    it honest against the generated FFI crate: every function it calls exists and gets the right number of arguments,
    every constant it names exists, every descriptor literal names exactly the struct's fields, in order."""
    sys_text = open(os.path.join(ROOT, "rust", "wgpu-b200-sys", "src", "lib.rs")).read()
    src = open(os.path.join(ROOT, "rust", "wgpu-b200", "src", "lib.rs")).read()
    src = re.sub(r"//[^\n]*", "", src)
    fns = {m.group(1): len(_split_top_level(m.group(2))) for m in re.finditer(r"pub fn (wgb_\w+)\((.*?)\)(?: ->|;)", sys_text)}
    consts = set(re.findall(r"pub const (WGB_\w+):", sys_text))
    structs = {m.group(1): re.findall(r"pub (\w+):", m.group(2)) for m in re.finditer(r"pub struct (wgb_\w+) \{\n(.*?)\n\}", sys_text, flags=re.S)}
    opaque = set(re.findall(r"pub struct (wgb_\w+_t) \{ _private", sys_text)) | set(re.findall(r"pub type (wgb_\w+) =", sys_text))
    called = 0
    for m in re.finditer(r"sys::(wgb_\w+)\(", src):
        name = m.group(1)
        assert name in fns, f"{name} is not declared by wgpu-b200-sys"
        args = _split_top_level(_balanced(src, m.end() - 1, "(", ")"))
        assert len(args) == fns[name], f"{name}: {len(args)} arguments, the C ABI takes {fns[name]}"
        called += 1
    assert called >= 45
    for name in set(re.findall(r"sys::(WGB_\w+)", src)):
        assert name in consts, f"{name} is not a constant of wgpu-b200-sys"
    for m in re.finditer(r"(?<!-> )sys::(wgb_\w+) \{", src):      # (not a function's return type)
        name = m.group(1)
        assert name in structs, f"{name} is not a descriptor struct"
        fields = [re.match(r"\s*(\w+)", p).group(1) for p in _split_top_level(_balanced(src, m.end() - 1, "{", "}")) if p.strip()]
        assert fields == structs[name], f"{name}: literal has fields {fields}, the struct has {structs[name]}"
    for name in set(re.findall(r"sys::(wgb_\w+)\b(?!\(| \{)", src)):
        assert name in opaque or name in structs or name in fns, f"sys::{name} does not exist"
    # every live trait of the reference backend has an impl here
    for trait in ("InstanceInterface", "AdapterInterface", "DeviceInterface", "QueueInterface", "BufferInterface", "BufferMappedRangeInterface",
                  "TextureInterface", "TextureViewInterface", "SamplerInterface", "ShaderModuleInterface", "BindGroupLayoutInterface",
                  "PipelineLayoutInterface", "BindGroupInterface", "RenderPipelineInterface", "CommandEncoderInterface",
                  "CommandBufferInterface", "RenderPassInterface", "SurfaceInterface", "SurfaceOutputDetailInterface"):
        assert re.search(r"impl %s for \w+" % trait, src), trait
