"""Run by tests/test_cusim_experiments.py in a subprocess (WGB_VARY_CACHE=1): three scenes on the software model with the
varying-cache experiment compiled into their pipelines, bit for bit against the oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import numpy as np  # noqa: E402

from tests.conftest import use_cusim  # noqa: E402

use_cusim()

from oracle import pyoracle  # noqa: E402
from wgpu_cpu_b200 import api, scenes as S  # noqa: E402
from wgpu_cpu_b200.render import SceneRenderer  # noqa: E402

dev, queue = api.instance().request_adapter().request_device(0)
for scene in (S.hello_mesh(256, 256), S.hello_texture(320, 180), S.synthetic_grid(384, 216, n=113, layers=4), S.multi_draw()):
    r = SceneRenderer(dev, queue, scene)
    assert "#define WGB_VARY_CACHE 1" in r.pipeline.get_source(), scene.name
    dev.poll(True, r.submit())
    ref = pyoracle.render(scene, want_coverage=False)
    assert np.array_equal(r.target.read(), ref.color), scene.name
    assert np.array_equal(r.depth_texture.read().view(np.uint32), ref.depth.view(np.uint32)), scene.name
    print("ok", scene.name)
