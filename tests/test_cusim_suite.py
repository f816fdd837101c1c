"""CPU tier: the GPU test-suite (tests marked `gpu`) run on tests/cusim -- a software model of the CUDA execution model
(fibers per thread, warp collectives, barriers, atomics, mbarrier, guard pages behind every device allocation) under
which the *unchanged* kernel source (wgb_raster.cuh, wgb_prelude.cuh), emitted shaders and host runtime (wgb_api.cpp)
execute on host cores.  This does not replace the B200 run -- scheduling, memory ordering and MUFU rounding are the
hardware's -- but on a machine without a GPU it checks the kernels themselves, not a restatement of them, against the
oracle, bit for bit.  The model is test infrastructure: the product library is not involved (and refuses to run
without a device, tests/test_c_abi.py::test_no_device_means_an_error_not_a_fallback).

The suite runs in a subprocess because the model build of the native library and the product build cannot share one
Python process (wgpu_cpu_b200.api holds one library handle)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# needs the real device: pinned host memory through torch.cuda
NEEDS_HARDWARE = ["tests/test_parity_gpu.py::test_pinned_uploads_on_the_copy_stream_are_ordered_with_rendering"]


def _xdist_args():
    """`-n <workers>` when pytest-xdist is there, else a serial run."""
    try:
        import xdist  # noqa: F401
    except ImportError:
        return []
    return ["-n", str(max(1, min(8, os.cpu_count() or 1)))]


@pytest.mark.cusim
def test_gpu_suite_on_the_software_model(tmp_path):
    from tests.cusim import build as cusim_build
    cusim_build.build()
    env = dict(os.environ, WGB_CUSIM="1", CUSIM_THREADS="2")
    env.setdefault("CUSIM_CACHE", str(tmp_path / "cache"))
    cmd = [sys.executable, "-m", "pytest", "tests", "-q", "-m", "gpu", "-x", "-p", "no:cacheprovider"] + _xdist_args()
    for t in NEEDS_HARDWARE:
        cmd += ["--deselect", t]
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=3000)
    tail = "\n".join(p.stdout.splitlines()[-40:])
    assert p.returncode == 0, f"GPU suite failed on the software model:\n{tail}\n{p.stderr[-2000:]}"
    # once more under a shuffled schedule (random start thread, direction and hold-backs per scheduler pass): the frame
    # must not depend on the order in which the threads of a CTA reach their barriers and collectives
    env["CUSIM_SCHED_SEED"] = "7"
    p = subprocess.run(cmd + ["-k", "not full_size"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=3000)   # (the 4K / 8K frames ran above)
    tail = "\n".join(p.stdout.splitlines()[-40:])
    assert p.returncode == 0, f"GPU suite failed on the software model under a shuffled schedule:\n{tail}\n{p.stderr[-2000:]}"
