"""bench.py's CPU rows: the secondary "all host cores" row renders slices of the draw on host threads and depth-composes
them.  The composition must be exactly the frame the single-threaded oracle renders (the oracle keeps no state between
calls; ties go to the earlier slice), or the row would time something else."""
import importlib.util
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from oracle import pyoracle
from wgpu_cpu_b200 import scenes as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_slices_on_threads_compose_to_the_single_threaded_frame():
    bench = _bench()
    scene = S.synthetic_grid(384, 216, n=113, layers=4)          # C3's shape: four overlapping layers, Less + depth write
    whole = pyoracle.render(scene, want_coverage=False)
    parts = bench.cpu_slices(scene, 7)
    assert sum(p.draws[0].count for p in parts) == scene.draws[0].count
    with ThreadPoolExecutor(len(parts)) as ex:
        frames = list(ex.map(lambda p: pyoracle.render(p, want_coverage=False), parts))
    color, depth = bench.cpu_compose(frames, 5)
    assert np.array_equal(color, whole.color)
    assert np.array_equal(depth.view(np.uint32), whole.depth.view(np.uint32))


def test_all_cores_row_reports_its_thread_count():
    bench = _bench()
    scene = S.synthetic_grid(192, 108, n=57, layers=2)
    row = bench.cpu_all_cores(scene, scene.num_primitives, threads=3)
    assert row["cores"] == 3 and row["value"] > 0 and "not the reference's behaviour" in row["note"]
    assert "unavailable" in bench.cpu_all_cores(S.procedural(64, 36), 2, threads=2)
