"""The compiled-language host (include/wgpu_b200.hpp + examples/hello_mesh.cpp: the reference's hello_mesh.rs flow in
C++ over the C ABI) builds, links against the library and -- without a GPU -- fails loudly instead of falling back."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLE = os.path.join(ROOT, "examples", "hello_mesh")


def test_example_builds_and_links():
    import sys
    sys.path.insert(0, ROOT)
    from wgpu_cpu_b200 import build
    build.build()
    for exe in (EXAMPLE, os.path.join(ROOT, "examples", "hello_texture"), os.path.join(ROOT, "examples", "hello_shader")):
        assert os.path.exists(exe)
        p = subprocess.run([exe], capture_output=True, text=True)
        assert p.returncode == 2 and "usage" in p.stderr


def test_example_reports_the_missing_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    wgsl = os.path.join(ROOT, "wgpu-cpu_b200", "shaders", "mesh_vertex_color.wgsl")
    p = subprocess.run([EXAMPLE, wgsl, os.devnull, os.devnull, os.devnull, "64", "64", "/tmp/wgb_example"], capture_output=True, text=True)
    assert p.returncode == 1 and "no CPU fallback" in p.stderr
