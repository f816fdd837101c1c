"""Asynchronous submissions (device.rs:436-462: submit returns an index at once; :258-289: poll waits).  Render passes whose
attachments are cleared on load are enqueued without waiting for their draws; `wgb_device_poll(wait)` -- or anything
else that looks at results -- settles them.  These tests keep several submissions in flight, with queue writes in
between, and check every frame against the oracle, including the cases where a pass in the middle of the queue has
to be run again (a work buffer overflowed) or raised an error."""
import math

import numpy as np
import pytest

from wgpu_cpu_b200 import scenes as S

pytestmark = pytest.mark.gpu


@pytest.fixture()
def gpu():
    from wgpu_cpu_b200 import api
    dev, queue = api.instance().request_adapter().request_device(0)
    return dev, queue


def _bunny_frames(n, size=(160, 120)):
    return [S.hello_texture(size[0], size[1], yaw=f * 2.0 * math.pi / n) for f in range(n)]


def test_frames_in_flight_with_uniform_writes_between_them(gpu):
    """A batch of frames like BASELINE's C5: one camera write + one submission per frame, each frame into its own
    target, ONE wait at the end."""
    from oracle import pyoracle
    from wgpu_cpu_b200.render import SceneRenderer
    dev, queue = gpu
    frames = _bunny_frames(6)
    targets = [dev.create_texture(frames[0].width, frames[0].height, frames[0].color_format) for _ in frames]
    r = SceneRenderer(dev, queue, frames[0], targets=targets)
    r.render()                                   # the draw shape's bin capacity is known from here on
    last = 0
    for k, sc in enumerate(frames):
        queue.write_buffer(r.resources[(0, 0)], 0, sc.bindings[(0, 0)][1])
        last = r.submit(r.encode(k))
    dev.poll(True, last)
    for k, sc in enumerate(frames):
        ref = pyoracle.render(sc, want_coverage=False)
        assert np.array_equal(targets[k].read(), ref.color), f"frame {k} differs"
    st = dev.last_pass_stats()
    assert st["replays"] == 0 and st["fragments"] > 0


def test_an_overflow_in_the_middle_of_the_queue_is_rerun_in_order(gpu, monkeypatch):
    """Frames A, B, A with different cameras: B piles the mesh into a corner of the frame, which overflows the bin capacity
    the first frame taught the device.  B's tile kernel poisons the queue, the third frame leaves its target alone, and
    the wait re-runs B and the third frame -- with the camera each of them was submitted with."""
    from oracle import pyoracle
    from wgpu_cpu_b200.render import SceneRenderer
    dev, queue = gpu
    a = S.hello_texture(256, 192)
    scale = np.diag([0.15, 0.15, 1.0, 1.0]).astype(np.float32)
    shift = np.eye(4, dtype=np.float32)
    shift[0, 3], shift[1, 3] = -0.8, 0.8
    cam_a = np.frombuffer(a.bindings[(0, 0)][1].tobytes(), dtype=np.float32).reshape(4, 4)       # column-major bytes
    cam_b = (shift @ scale @ cam_a.T).T.astype(np.float32)                                       # clip-space shrink towards a corner
    b = S.hello_texture(256, 192)
    b.bindings = dict(b.bindings)
    b.bindings[(0, 0)] = ("buffer", np.frombuffer(np.ascontiguousarray(cam_b).tobytes(), dtype=np.uint8).copy())
    seq = [a, b, a]
    targets = [dev.create_texture(a.width, a.height, a.color_format) for _ in seq]
    r = SceneRenderer(dev, queue, a, targets=targets)
    r.render()
    last = 0
    for k, sc in enumerate(seq):
        queue.write_buffer(r.resources[(0, 0)], 0, sc.bindings[(0, 0)][1])
        last = r.submit(r.encode(k))
    dev.poll(True, last)
    refs = [pyoracle.render(sc, want_coverage=False) for sc in seq]
    for k in range(3):
        assert np.array_equal(targets[k].read(), refs[k].color), f"frame {k} differs"
    assert not np.array_equal(refs[0].color, refs[1].color)


def test_an_error_in_the_middle_of_the_queue_surfaces_at_poll_and_later_submissions_still_run(gpu):
    from oracle import pyoracle
    from wgpu_cpu_b200 import api
    from wgpu_cpu_b200.render import SceneRenderer
    dev, queue = gpu
    good = S.hello_mesh(128, 96)
    bad = S.hello_mesh(128, 96)
    bad.index_data = bad.index_data.copy()
    bad.index_data[5] = 10_000_000                 # far outside the vertex buffer (vertex.rs:143-154: the slice panics)
    t_good = [dev.create_texture(128, 96, good.color_format) for _ in range(2)]
    rg = SceneRenderer(dev, queue, good, targets=t_good)
    rb = SceneRenderer(dev, queue, bad)
    rg.render()
    rg.submit(rg.encode(0))
    rb.submit()
    last = rg.submit(rg.encode(1))
    with pytest.raises(api.WgpuError):
        dev.poll(True, last)
    ref = pyoracle.render(good, want_coverage=False)
    assert np.array_equal(t_good[0].read(), ref.color)
    assert np.array_equal(t_good[1].read(), ref.color)


def test_synchronous_mode_gives_the_same_frames(monkeypatch):
    from oracle import pyoracle
    from wgpu_cpu_b200 import api
    from wgpu_cpu_b200.render import SceneRenderer
    monkeypatch.setenv("WGB_SYNC_SUBMIT", "1")
    dev, queue = api.instance().request_adapter().request_device(0)
    sc = S.hello_mesh(160, 120)
    r = SceneRenderer(dev, queue, sc)
    r.render()
    assert np.array_equal(r.read().color, pyoracle.render(sc, want_coverage=False).color)
