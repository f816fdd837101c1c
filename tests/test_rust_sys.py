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
