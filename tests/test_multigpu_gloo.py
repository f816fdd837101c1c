"""The N > 1 path on CPU: world_size-2 gloo process group, each rank owning one band of the frame
(wgpu_cpu_b200.multigpu), gathered to rank 0 with the same send/recv code bench.py runs over NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wgpu_cpu_b200.multigpu import band_rows, gather_bands


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, height, width, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        frame = torch.zeros(height, width, 4, dtype=torch.uint8)
        a, b = band_rows(height, rank, world)
        # stand-in for the tile stage: every rank writes only its own band
        frame[a:b] = torch.arange(a, b, dtype=torch.uint8).view(-1, 1, 1) + (rank + 1) * 40
        gather_bands(frame, rank, world, dst=0)
        dist.barrier()
        if rank == 0:
            np.save(out, frame.numpy())
    finally:
        dist.destroy_process_group()


def test_band_gather_world_size_2(tmp_path):
    height, width, world = 200, 16, 2
    out = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(world, _free_port(), height, width, out), nprocs=world, join=True)
    frame = np.load(out)
    for r in range(world):
        a, b = band_rows(height, r, world)
        expect = (np.arange(a, b, dtype=np.uint8) + (r + 1) * 40).reshape(-1, 1, 1)
        assert (frame[a:b] == expect).all(), f"band of rank {r} is wrong"


def test_bands_tile_the_frame():
    for h in (2160, 1080, 4320, 50):
        for n in (1, 2, 4, 8):
            rows = [band_rows(h, r, n) for r in range(n)]
            assert rows[0][0] == 0 and rows[-1][1] == h
            assert all(rows[i][1] == rows[i + 1][0] for i in range(n - 1))
            assert all(a % 32 == 0 or a == h for a, _ in rows)


def test_the_rows_left_over_go_to_the_middle_bands():
    """The outer bands hold the primitives that cross the top and bottom clip planes (C3 on eight GPUs: 7 195 clipped
    primitives on the top band against 1 563 on a middle one), so they are never the taller ones: band heights differ by
    at most one tile row, and they do not grow towards the edges."""
    for h in (2160, 1080, 4320, 50, 777):
        for n in (2, 3, 4, 7, 8):
            tiles = [(b - a + 31) // 32 for a, b in (band_rows(h, r, n) for r in range(n))]
            assert sum(tiles) == (h + 31) // 32 and max(tiles) - min(tiles) <= 1
            tall = [i for i, t in enumerate(tiles) if t == max(tiles)]
            if len(tall) < n:                  # the taller bands are one contiguous run in the middle
                assert tall == list(range(tall[0], tall[-1] + 1))
                assert abs(tall[0] - (n - 1 - tall[-1])) <= 1
    assert [(b - a + 31) // 32 for a, b in (band_rows(2160, r, 8) for r in range(8))] == [8, 8, 9, 9, 9, 9, 8, 8]


def _barrier_worker(rank, world, port, out):
    import time
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from wgpu_cpu_b200.multigpu import HostBarrier
        hb = HostBarrier(rank, world, f"test{port}")
        log = []
        for i in range(50):
            if rank == i % world:
                time.sleep(0.002)          # a straggler: nobody may pass before it arrives
            log.append(int(hb.slots[:, 0].min()))
            hb.wait()
            assert int(hb.slots[:, 0].min()) >= i + 1
        if rank == 0:
            np.save(out, np.asarray(log))
    finally:
        dist.destroy_process_group()


def test_host_barrier_world_size_2(tmp_path):
    out = str(tmp_path / "log.npy")
    mp.spawn(_barrier_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    log = np.load(out)
    assert (log >= np.arange(50)).all()


def _bench_dry_run(tmp_path, world, extra, config="c1"):
    """bench.py itself on the software model of tests/cusim (WGB_CUSIM=1: gloo instead of NCCL, "device" memory is host
    memory, the presenter's targets shared through a memfd): frame numbering, the presenter exchange with alternating
    targets, the sharded upload + all-gather of the e2e leg and the oracle digest of the assembled frame."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, WGB_CUSIM="1", CUSIM_THREADS="2", CUSIM_CACHE=str(tmp_path / "cache"))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), "bench.py", "--gpus", str(world), "--config", config, "--steps", "2", "--warmup", "1",
           "--no-cpu-baseline"] + extra
    p = subprocess.run(cmd, cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "rank 0 prints one JSON line"
    return json.loads(lines[0])


def test_bench_dry_run_two_ranks_peer_presenter(tmp_path):
    d = _bench_dry_run(tmp_path, 2, [])
    assert d["n_gpus"] == 2 and d["data"] == "software model dry run"
    assert d["parity"]["matches_oracle"] is True
    # every rank uploads half of the vertex and index bytes (+ the 16-byte tail and the uniforms): the scene once, not twice
    from wgpu_cpu_b200 import scenes as S
    sc = S.hello_mesh(512, 512)
    scene_bytes = sum(v.nbytes for v in sc.vertex_buffers) + sc.index_data.nbytes
    assert scene_bytes <= d["e2e"]["h2d_bytes_per_step"] <= scene_bytes + 2 * 64 + 64
    assert d["e2e"]["d2h_bytes_per_step"] == 512 * 512 * 4


def test_bench_dry_run_two_ranks_nccl_presenter(tmp_path):
    d = _bench_dry_run(tmp_path, 2, ["--present", "nccl", "--no-e2e"])
    assert d["parity"]["matches_oracle"] is True and "send/recv" in d["config"]["parallelism"]


def test_bench_dry_run_a_batch_of_frames_presented_round_robin(tmp_path):
    """A reduced C5 (4 frames per batch): with three ranks frame f is assembled on rank f mod 3 from the bands of all of
    them, every rank hashes the frames it presents, and the digest over the batch equals the one a single rank renders."""
    one = _bench_dry_run(tmp_path, 1, [], config="c5s")
    three = _bench_dry_run(tmp_path, 3, [], config="c5s")
    assert "rank f mod N" in three["config"]["parallelism"]
    assert three["parity"]["frame_sha256"] == one["parity"]["frame_sha256"]
    assert three["per_rank"] is not None and len(three["per_rank"]["tile_ms"]) == 3
