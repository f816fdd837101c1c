"""Tuning experiments that are compiled in only on request stay bit-exact: WGB_VARY_CACHE=1 (the tile kernel's shade
step reads the varyings the vertex kernel stored next to the post-transform positions instead of re-running the vertex
stage, wgb_raster.cuh wgb_prim_varyings) on the software model of tests/cusim, against the oracle."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_varying_cache_experiment_is_bit_exact_on_the_model(tmp_path):
    from tests.cusim import build as cusim_build
    cusim_build.build()
    env = dict(os.environ, WGB_VARY_CACHE="1", CUSIM_THREADS="2")
    env.setdefault("CUSIM_CACHE", str(tmp_path / "cache"))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "cusim", "vary_cache_check.py")], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert p.stdout.count("ok ") == 4
