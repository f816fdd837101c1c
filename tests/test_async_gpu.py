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


@pytest.mark.parametrize("camera_bytes", [64, 4096])
def test_frames_in_flight_with_uniform_writes_between_them(gpu, camera_bytes):
    """A batch of frames like BASELINE's C5: one camera write + one submission per frame, each frame into its own
    target, ONE wait at the end.  The tile kernel of frame k runs beside the geometry stage of frame k + 1 and reads the
    camera again when it shades: a small binding from the snapshot its pass took, a large one (the camera at the head
    of a 4 KB buffer) where it is -- the write for frame k + 1 then has to wait for it."""
    from oracle import pyoracle
    from wgpu_cpu_b200.render import SceneRenderer
    dev, queue = gpu
    frames = _bunny_frames(6)
    for sc in frames:
        cam = np.zeros(camera_bytes, dtype=np.uint8)
        cam[:64] = sc.bindings[(0, 0)][1]
        sc.bindings = dict(sc.bindings)
        sc.bindings[(0, 0)] = ("buffer", cam)
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


def _corner_camera(a):
    """The camera of scene `a` followed by a clip-space shrink towards a corner of the frame (piles the mesh into a few tiles)."""
    scale = np.diag([0.15, 0.15, 1.0, 1.0]).astype(np.float32)
    shift = np.eye(4, dtype=np.float32)
    shift[0, 3], shift[1, 3] = -0.8, 0.8
    cam_a = np.frombuffer(a.bindings[(0, 0)][1].tobytes(), dtype=np.float32).reshape(4, 4)
    cam_b = (shift @ scale @ cam_a.T).T.astype(np.float32)
    return np.frombuffer(np.ascontiguousarray(cam_b).tobytes(), dtype=np.uint8).copy()


def test_an_overflow_at_the_head_of_the_queue_sees_the_camera_written_before_it(gpu):
    """The pass that has to run again is the FIRST one in flight, and the camera it was submitted with was written while
    nothing was in flight; the write queued behind it has gone over that camera by the time the pass is run again."""
    from oracle import pyoracle
    from wgpu_cpu_b200.render import SceneRenderer
    dev, queue = gpu
    a = S.hello_texture(256, 192)
    b = S.hello_texture(256, 192)
    b.bindings = dict(b.bindings)
    b.bindings[(0, 0)] = ("buffer", _corner_camera(a))
    targets = [dev.create_texture(a.width, a.height, a.color_format) for _ in range(2)]
    r = SceneRenderer(dev, queue, a, targets=targets)
    r.render()
    dev.poll(True)
    queue.write_buffer(r.resources[(0, 0)], 0, b.bindings[(0, 0)][1])      # nothing in flight
    r.submit(r.encode(0))                                                   # overflows the bins frame `a` sized
    queue.write_buffer(r.resources[(0, 0)], 0, a.bindings[(0, 0)][1])      # behind it
    last = r.submit(r.encode(1))
    dev.poll(True, last)
    assert np.array_equal(targets[0].read(), pyoracle.render(b, want_coverage=False).color), "the re-run pass saw the wrong camera"
    assert np.array_equal(targets[1].read(), pyoracle.render(a, want_coverage=False).color)


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


def _pinned(nbytes):
    """Page-locked host bytes: torch's allocator on the device; on the software model every host pointer counts as
    page-locked when CUSIM_ALL_PINNED is set."""
    import os
    if os.environ.get("WGB_CUSIM"):
        return np.empty(nbytes, dtype=np.uint8) if os.environ.get("CUSIM_ALL_PINNED") else None
    import torch
    return torch.empty(nbytes, dtype=torch.uint8, pin_memory=True).numpy()


def test_a_read_back_that_is_not_waited_for_sees_its_frame_and_holds_back_the_next_writer(gpu):
    """wgb_texture_read_pinned_async: frame A is read back from target 0 while frame B renders into target 1 and frame C
    into target 0 again -- C's pass has to wait for the read-back, the host does not."""
    from oracle import pyoracle
    from wgpu_cpu_b200 import api
    from wgpu_cpu_b200.render import SceneRenderer
    dev, queue = gpu
    frames = _bunny_frames(3, size=(192, 128))
    targets = [dev.create_texture(192, 128, frames[0].color_format) for _ in range(2)]
    r = SceneRenderer(dev, queue, frames[0], targets=targets)
    r.render()
    nbytes = 192 * 128 * 4
    host = _pinned(nbytes)
    if host is None:
        with pytest.raises(api.WgpuError):
            targets[0].read_pinned_async(np.empty(nbytes, dtype=np.uint8))      # pageable memory is refused, not staged
        return
    queue.write_buffer(r.resources[(0, 0)], 0, frames[0].bindings[(0, 0)][1])
    dev.poll(True, r.submit(r.encode(0)))
    targets[0].read_pinned_async(host)
    last = 0
    for k in (1, 2):
        queue.write_buffer(r.resources[(0, 0)], 0, frames[k].bindings[(0, 0)][1])
        last = r.submit(r.encode(k))
    dev.wait_readbacks()
    refs = [pyoracle.render(sc, want_coverage=False).color for sc in frames]
    assert np.array_equal(host.reshape(refs[0].shape), refs[0]), "the read-back is not frame A"
    dev.poll(True, last)
    assert np.array_equal(targets[1].read(), refs[1])
    assert np.array_equal(targets[0].read(), refs[2])
    assert not np.array_equal(refs[0], refs[2])
