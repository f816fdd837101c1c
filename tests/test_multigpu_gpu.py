"""Two-GPU sort-first rendering with the peer-memory presenter: needs 2 CUDA devices (skipped otherwise).
Each rank is its own process (torch.multiprocessing), rank 1 maps rank 0's colour target over CUDA IPC and
its tile kernel stores its band straight into it; the assembled frame must equal the oracle's."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from wgpu_cpu_b200 import api, multigpu, scenes
        from wgpu_cpu_b200.render import SceneRenderer
        scene = scenes.hello_mesh(320, 200)
        dev, queue = api.instance().request_adapter().request_device(rank, band_rank=rank, band_count=world)
        own = dev.create_texture(scene.width, scene.height, scene.color_format) if rank == 0 else None
        target = multigpu.share_presenter_target(dev, own, rank, world, scene.width, scene.height, scene.color_format)
        r = SceneRenderer(dev, queue, scene, target=target)
        r.render()
        dist.barrier()
        if rank == 0:
            np.save(out, r.target.read())
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_gpu_peer_presenter_matches_oracle(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import pyoracle
    from wgpu_cpu_b200 import scenes
    out = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    ref = pyoracle.render(scenes.hello_mesh(320, 200))
    assert np.array_equal(np.load(out), ref.color)


def _model_worker(rank, world, port, out):
    """One process per "device" of the software model (tests/cusim): gloo for the process group, the colour target of
    rank 0 mapped into the other ranks through the model's cudaIpc* (a memfd), the tile kernels' bulk stores landing
    directly in it."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    from tests import conftest
    conftest.use_cusim()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from wgpu_cpu_b200 import api, multigpu, scenes
        from wgpu_cpu_b200.render import SceneRenderer
        scene = scenes.hello_mesh(320, 200)
        dev, queue = api.instance().request_adapter().request_device(rank, band_rank=rank, band_count=world)
        own = dev.create_texture(scene.width, scene.height, scene.color_format) if rank == 0 else None
        target = multigpu.share_presenter_target(dev, own, rank, world, scene.width, scene.height, scene.color_format)
        r = SceneRenderer(dev, queue, scene, target=target)
        r.render()
        dist.barrier()
        if rank == 0:
            np.save(out, r.target.read())
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_presenter_on_the_software_model(tmp_path, monkeypatch, world):
    """The sort-first bands and the peer-memory presenter with one process per rank, on the software model
    (WGB_CUSIM=1 only; the hardware version is the test above)."""
    if os.environ.get("WGB_CUSIM") != "1":
        pytest.skip("runs on the software model (WGB_CUSIM=1)")
    import torch.multiprocessing as mp
    from oracle import pyoracle
    from wgpu_cpu_b200 import scenes
    monkeypatch.setenv("CUSIM_IPC", "1")
    monkeypatch.setenv("CUSIM_DEVICES", str(world))
    out = str(tmp_path / "frame.npy")
    mp.spawn(_model_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    ref = pyoracle.render(scenes.hello_mesh(320, 200))
    assert np.array_equal(np.load(out), ref.color)
