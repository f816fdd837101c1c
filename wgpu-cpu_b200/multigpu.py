"""Sort-first multi-GPU partition: the framebuffer is split into `count` contiguous bands of
whole tile rows, one per rank; every rank holds the full scene, runs the geometry stage for all primitives
and the tile stage for its band only (per-pixel primitive order is preserved inside a band, so the result is
bit-identical to one GPU), and the colour bands are gathered to the presenting rank over NCCL/NVLink.

One process per GPU; torch.distributed supplies the process group (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

TILE_H = 32  # WGB_TILE_H


def band_rows(height: int, rank: int, count: int, tile_h: int = TILE_H):
    """First and one-past-last pixel row of band `rank` (must equal wgb_device_get_band_rows)."""
    tiles = (height + tile_h - 1) // tile_h
    q, rem = divmod(tiles, max(count, 1))
    lo = (max(count, 1) - rem) // 2          # the rows left over go to the middle bands (the outer ones hold the clipped primitives)
    extra_before = lambda k: min(max(k - lo, 0), rem)      # noqa: E731 -- taller bands among bands 0 .. k-1
    t0 = rank * q + extra_before(rank)
    t1 = (rank + 1) * q + extra_before(rank + 1)
    return min(t0 * tile_h, height), min(t1 * tile_h, height)


def gather_bands(frame, rank: int, world: int, dst: int = 0):
    """Gather every rank's band of `frame` (a [H, W, C] tensor that aliases the colour target) into rank
    `dst`'s copy.  Point-to-point send/recv of the band rows only: 4*W*H*(N-1)/N bytes per frame in total."""
    import torch.distributed as dist
    if world == 1:
        return
    height = frame.shape[0]
    if rank == dst:
        reqs = []
        for k in range(world):
            a, b = band_rows(height, k, world)
            if k != dst and b > a:
                reqs.append(dist.irecv(frame[a:b], src=k))
        for r in reqs:
            r.wait()
    else:
        a, b = band_rows(height, rank, world)
        if b > a:
            dist.send(frame[a:b].contiguous(), dst=dst)


def share_presenter_target(device, texture, rank: int, world: int, width: int, height: int, fmt: str, dst: int = 0):
    """Peer-memory presenter: rank `dst` exports its colour target over CUDA IPC, every other rank maps it and
    returns the mapped texture to use as its own colour attachment.  After this the tile kernels of all ranks
    store their bands directly into rank `dst`'s memory over NVLink; a barrier replaces the gather."""
    import torch.distributed as dist
    handles = [texture.export_ipc() if rank == dst else None]
    dist.broadcast_object_list(handles, src=dst)
    if rank == dst:
        return texture
    return device.import_texture_ipc(handles[0], width, height, fmt)


class HostBarrier:
    """Barrier between the rank processes of one node through a POSIX shared-memory page: every rank bumps
    its own 64-byte slot and spins until all slots reached the same count.  A few microseconds, against
    ~50-80 us for a NCCL barrier driven from Python; used between frames, where the ranks only have to agree
    that their tile kernels (which already stored the bands into the presenter's memory) have completed."""

    def __init__(self, rank: int, world: int, name: str):
        import mmap
        import os
        import torch.distributed as dist
        self.rank, self.world, self.count = rank, world, 0
        self.path = f"/dev/shm/wgb_barrier_{name}"
        if rank == 0:
            with open(self.path, "wb") as f:
                f.write(b"\0" * 64 * world)
        dist.barrier()
        self.fd = os.open(self.path, os.O_RDWR)
        self.mm = mmap.mmap(self.fd, 64 * world)
        import numpy as np
        self.slots = np.frombuffer(self.mm, dtype=np.int64).reshape(world, 8)
        dist.barrier()
        if rank == 0:
            os.unlink(self.path)

    def wait(self):
        self.count += 1
        self.slots[self.rank, 0] = self.count
        s = self.slots
        while int(s[:, 0].min()) < self.count:
            pass


def tensor_from_device_pointer(ptr: int, nbytes: int, device_index: int):
    """Zero-copy torch uint8 view of device memory owned by the backend (a texture's texel storage)."""
    import torch
    import os
    if os.environ.get("WGB_CUSIM") == "1":      # the software model of tests/cusim: "device" memory is host memory
        import ctypes
        import numpy as np
        return torch.from_numpy(np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(ptr)))

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}
    return torch.as_tensor(h, device=torch.device("cuda", device_index))
