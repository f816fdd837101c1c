"""TEST INFRASTRUCTURE: builds tests/cusim/_build/libwgpu_b200_sim.so -- the host runtime source of the product
(wgpu-cpu_b200/csrc/wgb_api.cpp, wgsl_emit.cpp, unchanged) compiled against the stand-in CUDA headers of
tests/cusim/include and linked with cusim_host.cpp instead of the CUDA runtime.  See cusim_device.h."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "wgpu-cpu_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libwgpu_b200_sim.so")


def build(force: bool = False) -> str:
    """CUSIM_HOST_COVERAGE=1: an -O0 --coverage build (libwgpu_b200_sim_cov.so; the .gcno / .gcda files land in _build/),
    for measuring which lines of the host runtime and the WGSL emitter the tests execute."""
    import importlib
    import sys
    coverage = os.environ.get("CUSIM_HOST_COVERAGE") == "1"
    # CUSIM_ASAN=1: the host runtime under AddressSanitizer (the fibers of the device half are not instrumented).  Python
    # is not, so the sanitizer runtime has to be preloaded together with libstdc++ (for its __cxa_throw interceptor):
    #   CUSIM_ASAN=1 WGB_CUSIM=1 ASAN_OPTIONS=detect_leaks=0 \
    #   LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libstdc++.so.6)" python -m pytest tests -m gpu -q
    asan = os.environ.get("CUSIM_ASAN") == "1"
    lib = os.path.join(OUT, "libwgpu_b200_sim_cov.so") if coverage else os.path.join(OUT, "libwgpu_b200_sim_asan.so") if asan else LIB
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    product = importlib.import_module("wgpu_cpu_b200.build")
    emb = product._embed()                                   # the device sources as text, exactly as the product embeds them
    srcs = [os.path.join(CSRC, "wgb_api.cpp"), os.path.join(CSRC, "wgsl_emit.cpp"), emb, os.path.join(HERE, "cusim_host.cpp")]
    deps = srcs + [os.path.join(HERE, "include", f) for f in os.listdir(os.path.join(HERE, "include"))] + [
        os.path.join(CSRC, "wgb_shared.h"), os.path.join(ROOT, "include", "wgpu_b200.h"), os.path.abspath(__file__)]
    if not force and os.path.exists(lib) and all(os.path.getmtime(d) <= os.path.getmtime(lib) for d in deps):
        return lib
    os.makedirs(OUT, exist_ok=True)
    cmd = ["g++", *(["-O0", "--coverage"] if coverage else ["-O1", "-fsanitize=address", "-fno-omit-frame-pointer"] if asan else ["-O1"]), "-g", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-w",
           f"-I{os.path.join(HERE, 'include')}", f'-DCUSIM_DIR="{HERE}"', *srcs, "-o", lib, "-ldl", "-lpthread"]
    subprocess.check_call(cmd, cwd=OUT)
    return lib


def build_example(name: str = "hello_mesh") -> str:
    """examples/<name>.cpp (the C++ host over the C ABI) linked against the model build of the library."""
    lib = build()
    src = os.path.join(ROOT, "examples", name + ".cpp")
    exe = os.path.join(OUT, name)
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(lib)):
        subprocess.check_call(["g++", "-std=c++17", "-O1", f"-I{os.path.join(ROOT, 'include')}", src, "-o", exe,
                               f"-L{OUT}", "-lwgpu_b200_sim", f"-Wl,-rpath,{OUT}"])
    return exe


if __name__ == "__main__":
    print(build(force=True))
