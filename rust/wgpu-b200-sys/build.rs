// WGPU_B200_LIB_DIR = the directory that holds libwgpu_b200.so (python -c "import __graft_entry__ as g; g.build()")
fn main() {
    if let Ok(dir) = std::env::var("WGPU_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=wgpu_b200");
    println!("cargo:rerun-if-env-changed=WGPU_B200_LIB_DIR");
}
