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
