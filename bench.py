#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 render-pass draw path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config c3|c1|c2|c4]

A step is one render pass (clear + draw + store) of the workload BASELINE.json's metric is quoted on:
config C3, the synthetic ~10M-triangle mesh with vertex colours + depth at 3840x2160.
One JSON line is printed by rank 0:
  value       Mtri/s, whole job, inputs resident in HBM, K passes between device syncs
  e2e         the same metric through the public API with HOST inputs: every step uploads the vertex,
              index and uniform data from pinned host memory and reads the colour target back
  roofline    the dominant kernel (the tile kernel) against the measured HBM peak
  cpu_baseline the CPU oracle (a C++ port of the reference's algorithm, 1 thread like the reference's
              engine thread) timed on a bounded sample of the same workload on this box's host cores
With --gpus N > 1 (launched under torchrun) the framebuffer is split sort-first into N bands of tile
rows, every rank runs the geometry stage for all triangles and the tile stage for its band, and the
colour bands are gathered to rank 0 with NCCL inside the timed region ("strong" scaling: the frame is
fixed).  --impl reference times the CPU port alone.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "c1": ("hello_mesh teapot 512x512 (C1)", lambda S: S.hello_mesh(512, 512)),
    "c2": ("hello_texture bunny 1920x1080 (C2)", lambda S: S.hello_texture(1920, 1080)),
    "c3": ("synthetic 10M-triangle mesh, vertex colours + depth, 3840x2160 (C3)", lambda S: S.synthetic_grid(3840, 2160, 1119, 4)),
    "c3s": ("synthetic 1M-triangle mesh 3840x2160 (reduced C3, smoke only)", lambda S: S.synthetic_grid(3840, 2160, 354, 4)),
    "c4": ("full-screen 64-iteration fragment shader 7680x4320 (C4)", lambda S: S.procedural(7680, 4320)),
    "c5": ("batch of 64 frames of the textured bunny at 3840x2160, camera yaw = frame*tau/64 (C5); one step = 64 passes",
           lambda S: S.hello_texture(3840, 2160)),
    "c5s": ("batch of 4 frames of the textured bunny at 320x256 (reduced C5, smoke only); one step = 4 passes", lambda S: S.hello_texture(320, 256)),
}
BATCH_FRAMES = {"c5": 64, "c5s": 4}      # configs whose step is a batch of independent frames


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md clocks line).  Samples through NVML every
    few milliseconds from a thread (the timed region of a millisecond-frame benchmark is shorter than nvidia-smi's
    start-up time); falls back to an `nvidia-smi -lms` subprocess when the NVML module is unavailable.  The sampler
    runs from before the warm-up; `mark()` opens and `stop()` closes the window whose samples are reported."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.nvml = index, [], None, None     # rows: (time, sm_mhz, max_mhz, {reasons})
        self.t_mark, self.running = None, False

    def start(self):
        self.running = True
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = (pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index))
            threading.Thread(target=self._poll_nvml, daemon=True).start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read_smi, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        nv, h = self.nvml
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        while self.running:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.rows.append((time.perf_counter(), sm, mx, {name for bit, name in self.REASONS if bits & bit}))
            except Exception:
                pass
            time.sleep(0.004)

    def _read_smi(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            if len(r) >= 9 and r[1].replace(".", "").isdigit():
                reasons = {name for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9])
                           if v.lower().startswith("active")}
                self.rows.append((time.perf_counter(), float(r[1]), float(r[2]) if r[2].replace(".", "").isdigit() else None, reasons))

    def mark(self):
        self.t_mark = time.perf_counter()

    def stop(self) -> dict:
        t_end = time.perf_counter()
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML and no nvidia-smi"], "samples": 0}
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
        self.running = False
        t0 = self.t_mark if self.t_mark is not None else 0.0
        inside = [r for r in self.rows if t0 <= r[0] <= t_end]
        note = None
        if not inside and self.rows:        # a window shorter than the sampling period: the samples next to it (still under load: warm-up precedes it)
            inside = sorted(self.rows, key=lambda r: min(abs(r[0] - t0), abs(r[0] - t_end)))[:2]
            note = "timed window shorter than the sampling period: nearest samples"
        sm = [r[1] for r in inside]
        mx = [r[2] for r in inside if r[2] is not None]
        reasons = set().union(*[r[3] for r in inside]) if inside else set()
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}
        if note:
            out["note"] = note
        return out


def cpu_baseline(scene, budget_s: float = 25.0) -> dict:
    """Time the CPU oracle (1 thread) on a bounded sample of the workload: the whole frame if it fits the
    budget, otherwise a prefix of the draw's triangles scaled from a short probe."""
    import copy
    from oracle import pyoracle
    from wgpu_cpu_b200 import scenes as S
    d = scene.draws[0]
    probe = copy.copy(scene)
    n_probe = min(d.count, 300_000 // 3 * 3) if scene.topology == "triangle-list" else d.count
    probe.draws = [S.Draw(d.indexed, d.first, n_probe, d.base_vertex, d.first_instance, d.instance_count)]
    t = time.perf_counter()
    pyoracle.render(probe, want_coverage=False)
    dt = time.perf_counter() - t
    count = d.count
    if n_probe < d.count:
        count = int(min(d.count, max(n_probe, n_probe * budget_s / max(dt, 1e-6)))) // 3 * 3
    sample = copy.copy(scene)
    sample.draws = [S.Draw(d.indexed, d.first, count, d.base_vertex, d.first_instance, d.instance_count)]
    t = time.perf_counter()
    fr = pyoracle.render(sample, want_coverage=False)
    dt = time.perf_counter() - t
    tris = fr.stats["primitives_assembled"]
    out = {"value": tris / dt / 1e6, "unit": "Mtri/s", "cores": 1, "kind": "port",
           "sample": f"first {tris} of {scene.num_primitives} triangles of the same frame, 1 pass, {dt:.1f} s "
                     f"(C++ oracle port of the reference algorithm, g++ -O2, single thread like the reference's engine thread)",
           "fragments_mpix_s": fr.stats["fragments_shaded"] / dt / 1e6, "seconds": dt}
    try:
        out["all_cores"] = cpu_all_cores(sample, tris)
    except Exception as e:          # the secondary row must not cost the line its primary one
        out["all_cores"] = {"unavailable": str(e)[:200]}
    return out


def cpu_slices(scene, parts: int):
    """`scene` once per contiguous slice of its draw's triangles (sort-last on host threads)."""
    import copy
    from wgpu_cpu_b200 import scenes as S
    d = scene.draws[0]
    ntri = d.count // 3
    parts = max(1, min(parts, ntri))
    edges = [ntri * k // parts for k in range(parts + 1)]
    out = []
    for k in range(parts):
        s = copy.copy(scene)
        s.draws = [S.Draw(d.indexed, d.first + 3 * edges[k], 3 * (edges[k + 1] - edges[k]), d.base_vertex, d.first_instance, d.instance_count)]
        out.append(s)
    return out


def cpu_compose(frames, bands: int):
    """Depth-compose the slices' frames: the smallest depth wins and a tie goes to the earlier slice -- with Less + depth
    write and no blending that is the serial result (a later fragment replaces a texel only if strictly nearer)."""
    from concurrent.futures import ThreadPoolExecutor
    h = frames[0].depth.shape[0]
    color = np.empty_like(frames[0].color)
    depth = np.empty_like(frames[0].depth)
    edges = [h * b // bands for b in range(bands + 1)]

    def band(b):
        y0, y1 = edges[b], edges[b + 1]
        if y0 == y1:
            return
        dz = np.stack([f.depth[y0:y1] for f in frames])
        win = np.argmin(dz, axis=0)                      # first occurrence of the minimum = the earlier slice
        depth[y0:y1] = np.take_along_axis(dz, win[None], axis=0)[0]
        dc = np.stack([f.color[y0:y1] for f in frames])
        color[y0:y1] = np.take_along_axis(dc, win[None, :, :, None], axis=0)[0]

    with ThreadPoolExecutor(bands) as ex:
        list(ex.map(band, range(bands)))
    return color, depth


def cpu_all_cores(sample, tris: int, threads: int = 0) -> dict:
    """Secondary CPU row (optional): the oracle over slices of the draw on all host cores, depth-composed.
    NOT the reference's behaviour -- its render path is one engine thread (engine.rs:15-24) -- so the primary row stays
    single-threaded.  (Row bands, the other split, gain nothing at C3: every band replays the per-triangle work.)"""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle
    if not (sample.topology == "triangle-list" and sample.has_depth and sample.depth_write and sample.depth_compare == "less"
            and not getattr(sample, "blend", None) and len(sample.draws) == 1 and sample.draws[0].instance_count == 1):
        return {"unavailable": "the depth composition of slices is the serial frame only for one triangle list under Less + depth write"}
    threads = threads or max(1, min(os.cpu_count() or 1, 32))
    parts = cpu_slices(sample, threads)
    t = time.perf_counter()
    with ThreadPoolExecutor(len(parts)) as ex:          # ctypes releases the GIL inside the oracle, which keeps no global state
        frames = list(ex.map(lambda s: pyoracle.render(s, want_coverage=False), parts))
    cpu_compose(frames, len(parts))
    dt = time.perf_counter() - t
    return {"value": tris / dt / 1e6, "unit": "Mtri/s", "cores": len(parts), "seconds": dt,
            "note": "slices of the draw over host threads + depth composition: not the reference's behaviour (single engine thread)"}


def run_reference(args, scene, workload):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each step is a bounded sample sized so that warmup + steps stay within a few minutes
    steps, warm = max(1, args.steps), max(0, args.warmup)
    budget = max(2.0, min(25.0, 150.0 / (steps + warm)))
    vals, frs, last = [], [], None
    for i in range(warm + steps):
        last = cpu_baseline(scene, budget)
        if i >= warm:
            vals.append(last["value"])
            frs.append(last["fragments_mpix_s"])
    v = float(np.mean(vals))
    cb = dict(last)
    cb["value"] = v
    line = {"impl": "reference", "metric": "Mtri/s", "value": v, "unit": "Mtri/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": last["seconds"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": workload}, "fragments_mpix_s": float(np.mean(frs)),
            "cpu_baseline": cb, "e2e": {"value": v, "unit": "Mtri/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="c3")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--present", default="peer", choices=["peer", "nccl"],
                    help="N>1: how colour bands reach rank 0: tile kernels store into rank 0's target over NVLink peer memory "
                         "(fused, default) or a separate NCCL gather")
    args = ap.parse_args()

    from wgpu_cpu_b200 import scenes as S
    workload, make = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            run_reference(args, make(S), workload)
        return

    import torch
    import torch.distributed as dist
    from wgpu_cpu_b200 import api
    from wgpu_cpu_b200.render import SceneRenderer

    # WGB_CUSIM=1: a dry run of this script's logic (frame numbering, presenter exchange, parity digests) on the software
    # model of tests/cusim, for development on a machine without a GPU -- test infrastructure, its numbers mean nothing
    # and its line says so ("data": "software model dry run")
    model = os.environ.get("WGB_CUSIM") == "1"
    if model:
        os.environ.setdefault("CUSIM_ALL_PINNED", "1")
        os.environ.setdefault("CUSIM_IPC", "1")
        os.environ.setdefault("CUSIM_DEVICES", str(max(world, 1)))
        from tests.cusim import build as cusim_build
        api.LIB_PATH = cusim_build.build()
    elif not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the B200 backend has no CPU fallback)")
    tdev = "cpu" if model else "cuda"
    sync = (lambda: None) if model else torch.cuda.synchronize
    if not model:
        torch.cuda.set_device(local_rank)
    if world > 1:
        if model:
            dist.init_process_group("gloo")
        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    scene = make(S)
    W, H = scene.width, scene.height
    band_rank, band_count = rank, world
    if world == 1 and os.environ.get("WGB_BENCH_BAND"):      # profiling aid: one rank's share of an N-way partition on one GPU
        band_rank, band_count = (int(v) for v in os.environ["WGB_BENCH_BAND"].split("/"))
    dev, queue = api.instance().request_adapter().request_device(local_rank, band_rank=band_rank, band_count=band_count)
    from wgpu_cpu_b200 import multigpu
    use_emitted = os.environ.get("WGB_USE_EMITTED", "0") == "1"
    # N > 1: two presenter targets, used alternately, so that a rank that is a frame ahead never stores into the frame
    # rank 0 is still reading (one barrier per frame is then enough)
    targets = None
    batch = BATCH_FRAMES.get(args.config)
    n_present = (batch if batch else 2) if world > 1 else 1

    def presenter_of(frame):
        """The rank a frame is assembled on.  A batch of independent frames is presented round-robin -- frame f on rank
        f mod N -- so that the bands of one frame converge on one GPU and those of the next on another (all of them
        into rank 0: its NVLink ingest, 7/8 of every frame, bounds the tile kernels of the whole batch)."""
        return (frame % n_present) % world if batch and args.present == "peer" else 0

    if world > 1 and args.present == "peer":
        targets = []
        for i in range(n_present):
            dst = presenter_of(i)
            own = dev.create_texture(W, H, scene.color_format) if rank == dst else None
            targets.append(multigpu.share_presenter_target(dev, own, rank, world, W, H, scene.color_format, dst=dst))
    elif world > 1:
        targets = [dev.create_texture(W, H, scene.color_format) for _ in range(n_present)]
    r = SceneRenderer(dev, queue, scene, use_emitted=use_emitted, targets=targets)
    warm = max(args.warmup, 3)

    # presenter: every rank's colour band -> rank 0
    row0, row1 = dev.band_rows(H)
    assert (row0, row1) == multigpu.band_rows(H, band_rank, band_count)
    frame_ts = None
    host_barrier = multigpu.HostBarrier(rank, world, os.environ.get("MASTER_PORT", "0")) if world > 1 and args.present == "peer" else None
    if world > 1 and args.present == "nccl":
        frame_ts = []
        for t in r.targets:
            ptr, nbytes = t.device_pointer()
            frame_ts.append(multigpu.tensor_from_device_pointer(ptr, nbytes, local_rank).view(H, W, 4))

    def gather(which=0):
        if world == 1:
            return
        if args.present == "peer":
            # the tile kernels already stored every band into rank 0's target over NVLink; the render pass has
            # completed on each rank's stream (poll(Wait)), so a barrier between the rank processes (shared-memory
            # page, no collective) makes the frame visible to the presenter
            host_barrier.wait()
            return
        # the render pass has completed on the backend's stream (poll(Wait)); NCCL runs on torch's stream
        multigpu.gather_bands(frame_ts[which % n_present], rank, world, dst=0)
        sync()

    def barrier():
        if world > 1:
            dist.barrier()
        sync()

    # C5: a step is a batch of 64 frames, each with its own camera (a 64-byte uniform update per frame)
    cameras = None
    if batch:
        import math
        cameras = [S.hello_texture(W, H, yaw=f * 2.0 * math.pi / batch).bindings[(0, 0)][1] for f in range(batch)]
    passes_per_step = len(cameras) if cameras else 1

    # The timed span mirrors the reference's own (render_pass/mod.rs:346-392: State::new -> load -> draws -> store, i.e.
    # execution, not recording -- SURVEY 8d): command buffers are recorded before the timed region, one per step, and a
    # step is submit + poll(Wait) (+ the presenter exchange)
    recorded = {}
    frame_no = [0]          # frames rendered so far: frame k goes to presenter target k % n_present

    def step():
        """One step, waited for: submit + poll(Wait) + presenter exchange; returns the pass statistics (a C5 batch: their sum)."""
        k = frame_no[0]
        if cameras is None:
            frame_no[0] += 1
            st = r.render(recorded.pop(k, None) or r.encode(k))
            gather(k)
            return st
        acc = None
        for cam in cameras:
            k = frame_no[0]
            frame_no[0] += 1
            queue.write_buffer(r.resources[(0, 0)], 0, cam)
            st = r.render(r.encode(k))
            gather(k)
            if acc is None:
                acc = dict(st)
            else:
                for key in ("primitives", "fragments", "shaded", "bin_pairs", "hiz_culled", "big_primitives", "clipped_primitives", "clip_records",
                            "kernel_launches", "replays", "geometry_ms", "tile_ms", "total_ms"):
                    acc[key] += st[key]
        return acc

    def run_steps(n):
        """The timed body: n steps with the submissions kept in flight (submit returns at once, device.rs:436-462): step
        k + 1 is enqueued before step k is waited for, so the device never idles on the host.  A C5 step enqueues its 64
        frames -- camera write + submission each, every frame into its own presenter target -- and waits once."""
        if cameras is None:
            prev = None
            for _ in range(n):
                k = frame_no[0]
                frame_no[0] += 1
                idx = r.submit(recorded.pop(k, None) or r.encode(k))
                if prev is not None:
                    dev.poll(True, prev[0])
                    gather(prev[1])
                prev = (idx, k)
            dev.poll(True, prev[0])
            gather(prev[1])
            return
        for _ in range(n):
            first, idx = frame_no[0], 0
            frame_no[0] += len(cameras)
            # The frames of a batch are independent (own camera, own target), so a rank is free to render them in any
            # order: rank r starts at frame r.  At any moment the N ranks are then on N different frames, whose bands
            # converge on N different presenters (frame f assembles on rank f mod N) -- started together on frame 0
            # they all store into one GPU at once, and its NVLink ingest (7/8 of a frame per pass) is what the tile
            # kernels of the whole batch wait for.
            shift = rank % len(cameras) if world > 1 and args.present == "peer" else 0
            for j in range(len(cameras)):
                f = (j + shift) % len(cameras)
                k = first + f
                queue.write_buffer(r.resources[(0, 0)], 0, cameras[f])
                idx = r.submit(recorded.pop(k, None) or r.encode(k))
            dev.poll(True, idx)
            if world > 1 and args.present == "nccl":
                for k in range(first, frame_no[0]):
                    gather(k)
            else:
                gather(first)          # one exchange per batch: every frame of the batch went to its own presenter target

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(warm):
        step()
    run_steps(2)          # and two steps the way the timed ones run (in flight together: the second set of work buffers, the tile stream)
    for k in range(frame_no[0], frame_no[0] + args.steps * passes_per_step):      # recording is outside the timed span, as for the reference's pass timer
        recorded[k] = r.encode(k)

    # ---- timed: K passes, inputs resident in HBM ----
    barrier()
    sampler.mark()
    event_ms = None
    try:
        dev.timer_begin()              # CUDA event on the render stream (the stream every kernel of the path is launched on)
    except Exception:                  # noqa: BLE001 -- the wall clock below still brackets the region
        pass
    t0 = time.perf_counter()
    run_steps(args.steps)
    try:
        event_ms = float(dev.timer_end())   # second event on the same stream, synchronised
    except Exception:                  # noqa: BLE001
        event_ms = None
    barrier()
    wall_dt = time.perf_counter() - t0
    clocks = sampler.stop()
    # the timed region is measured on the device (CUDA events on the launching stream); the wall clock around the same
    # region, bracketed by barrier + synchronize on both sides, is reported next to it and is used only if the events fail
    dt = event_ms * 1e-3 if event_ms and event_ms > 0 else wall_dt
    if world > 1:
        tmax = torch.tensor([dt], device=tdev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dt = float(tmax.item())

    # per-kernel device times (CUDA events around the geometry stage and the tile kernel of every pass) and the pass
    # counters: from a few more steps that are waited for one by one, outside the timed region
    stats = [step() for _ in range(3)]
    prims = scene.num_primitives * passes_per_step
    ms_step = dt / args.steps * 1e3
    tile_ms = float(np.mean([s["tile_ms"] for s in stats]))
    geom_ms = float(np.mean([s["geometry_ms"] for s in stats]))
    dev_ms = float(np.mean([s["total_ms"] for s in stats]))
    last = stats[-1]
    # every rank's share (sort-first: the geometry stage is replicated, the tile work follows the scene): rank 0 reports them all
    per_rank = None
    if world > 1:
        mine = torch.tensor([geom_ms, tile_ms, dev_ms, float(last["bin_pairs"]), float(last["fragments"])], device=tdev, dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"geometry_ms": [round(float(t[0]), 4) for t in allr], "tile_ms": [round(float(t[1]), 4) for t in allr],
                    "device_ms": [round(float(t[2]), 4) for t in allr], "bin_pairs": [int(t[3]) for t in allr], "fragments": [int(t[4]) for t in allr]}

    # ---- e2e: host buffers in, colour target out, every step ----
    e2e = None
    if not args.no_e2e and cameras is None:
        # every step uploads one full input set from pinned host memory, renders it and reads the frame back.  At N > 1
        # each rank uploads 1/N of the vertex and index bytes over its own PCIe link and the ranks all-gather the
        # slices over NVLink (NCCL, in place in the backend's buffers), so the host feeds the scene once, not N times.
        # three resident input sets at N > 1 (upload of k+2, all-gather of k+1 and rendering of k run side by side), two at N = 1
        n_sets = 3 if world > 1 else 2
        sets = [r] + [SceneRenderer(dev, queue, scene, use_emitted=use_emitted, targets=targets) for _ in range(n_sets - 1)]

        def pin(a):
            t = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=not model)
            t.numpy()[:] = a.view(np.uint8).reshape(-1)
            return t

        big = [("vb", i, pin(vb)) for i, vb in enumerate(scene.vertex_buffers)]
        if scene.index_data is not None:
            big.append(("ib", 0, pin(scene.index_data)))
        uni = [(key, pin(res[1])) for key, res in scene.bindings.items() if res[0] == "buffer"]

        def buf_of(rr, kind, i):
            return rr.vertex_buffers[i] if kind == "vb" else rr.index_buffer

        def chunk(n):           # bytes per rank in the all-gathered part (16-byte granules); the tail is uploaded by every rank
            return (n // world) & ~15 if world > 1 else n

        alias = {}
        if world > 1:
            for si, rr in enumerate(sets):
                for kind, i, t in big:
                    ptr, nbytes = buf_of(rr, kind, i).device_pointer()
                    alias[(si, kind, i)] = multigpu.tensor_from_device_pointer(ptr, nbytes, local_rank)
        h2d_rank = sum(chunk(t.numel()) + (t.numel() - chunk(t.numel()) * world if world > 1 else 0) for _, _, t in big) + sum(t.numel() for _, t in uni)
        d2h = W * H * 4
        frame_host = [torch.empty(d2h, dtype=torch.uint8, pin_memory=not model) for _ in range(2)]

        def upload(si):
            rr = sets[si]
            for kind, i, t in big:
                n, c = t.numel(), chunk(t.numel())
                a = t.numpy()
                if c:
                    queue.write_buffer_pinned_async(buf_of(rr, kind, i), rank * c if world > 1 else 0, a[rank * c:(rank + 1) * c] if world > 1 else a)
                if world > 1 and n > c * world:
                    queue.write_buffer_pinned_async(buf_of(rr, kind, i), c * world, a[c * world:])
            for key, t in uni:
                queue.write_buffer_pinned_async(rr.resources[key], 0, t.numpy())

        def exchange_start(si):
            """All-gather of the ranks' slices of input set `si`, in place in the backend's buffers, on torch's stream; not
            waited for here.  The caller has waited for this rank's slice (queue.wait_uploads)."""
            for kind, i, t in big:
                c = chunk(t.numel())
                if c:
                    full = alias[(si, kind, i)]
                    dist.all_gather_into_tensor(full[:c * world], full[rank * c:(rank + 1) * c])

        e2e_k = [0]

        def e2e_step():
            """Step k renders input set k % n_sets and starts the read-back of its frame.  Beside it: the upload of the
            set after next on the copy stream, the all-gather of the next set on torch's stream (N > 1), and the
            read-back of the previous frame on the read-back stream (the presenter targets alternate)."""
            k = e2e_k[0]
            e2e_k[0] += 1
            cur = k % n_sets
            if world > 1:
                nxt = (k + 1) % n_sets
                queue.wait_uploads()                      # this rank's slice of set k+1, started a step ago
                exchange_start(nxt)
                upload((k + 2) % n_sets)                  # last used by step k-1, which has been waited for
            else:
                upload((k + 1) % n_sets)
            f = frame_no[0]
            frame_no[0] += 1
            dev.poll(True, sets[cur].submit(sets[cur].encode(f)))
            if rank == 0:
                dev.wait_readbacks()                      # frame f-1 is on the host: after the exchange below the ranks may go on to frame f+1, which overwrites its target
            gather(f)
            if rank == 0:
                sets[cur].targets[f % n_present].read_pinned_async(frame_host[f % 2].numpy())
            if world > 1:
                sync()                                    # the all-gather of set k+1 has completed

        upload(0)
        if world > 1:
            queue.wait_uploads()
            exchange_start(0)
            sync()
            upload(1)
        for _ in range(3):
            e2e_step()
        n_e2e = max(4, min(args.steps, 10)) & ~1
        queue.wait_uploads()
        dev.wait_readbacks()
        barrier()
        t1 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        queue.wait_uploads()
        dev.poll(True)
        dev.wait_readbacks()
        barrier()
        de = time.perf_counter() - t1
        if world > 1:
            tmax = torch.tensor([de], device=tdev)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            de = float(tmax.item())
            tsum = torch.tensor([float(h2d_rank)], device=tdev, dtype=torch.float64)
            dist.all_reduce(tsum)
            h2d_total = int(tsum.item())
        else:
            h2d_total = int(h2d_rank)
        e2e = {"value": prims * n_e2e / de / 1e6, "unit": "Mtri/s", "h2d_bytes_per_step": h2d_total, "d2h_bytes_per_step": int(d2h),
               "ms_per_step": de / n_e2e * 1e3, "steps": n_e2e,
               "pipelining": "while step k renders, the inputs of a later step upload on the copy stream and frame k-1 is read back on the read-back stream" +
                             ("; every rank uploads 1/N of the vertex and index bytes of step k+2 while NCCL all-gathers those of step k+1 over NVLink (three resident input sets)" if world > 1 else " (two resident input sets)")}

    # ---- parity: the frame this run assembles on rank 0 against the oracle's digest (tests/golden, tools/make_bench_golden.py) ----
    parity = None
    if not (world == 1 and os.environ.get("WGB_BENCH_BAND")):
        import hashlib
        golden = {}
        try:
            golden = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_frames_sha256.json"))).get(args.config, {})
        except Exception:      # noqa: BLE001
            pass
        frames = None
        if cameras is None:
            k = frame_no[0]
            step()
            barrier()
            digest = hashlib.sha256(np.ascontiguousarray(r.targets[k % n_present].read()).tobytes()).hexdigest() if rank == 0 else None
            want = golden.get("color")
        elif world > 1 and args.present == "peer":
            # C5: one more batch rendered the way the timed ones are (all frames in flight, every rank in its own frame
            # order, one wait), then every frame of it is hashed on the rank it assembled on
            first = frame_no[0]
            run_steps(1)
            barrier()
            frames = [hashlib.sha256(np.ascontiguousarray(r.targets[(first + f) % n_present].read()).tobytes()).hexdigest()
                      if rank == presenter_of(first + f) else None for f in range(len(cameras))]
            barrier()
        else:       # C5 on one GPU (one target) or gathered by NCCL: every frame of one more batch, one at a time
            frames = []
            for cam in cameras:
                k = frame_no[0]
                frame_no[0] += 1
                queue.write_buffer(r.resources[(0, 0)], 0, cam)
                r.render(r.encode(k))
                gather(k)
                frames.append(hashlib.sha256(np.ascontiguousarray(r.targets[k % n_present].read()).tobytes()).hexdigest()
                              if rank == presenter_of(k) else None)
                barrier()
        if frames is not None:
            if world > 1:           # every rank hashed the frames it presents
                every = [None] * world
                dist.all_gather_object(every, frames)
                frames = [next(d[i] for d in every if d[i] is not None) for i in range(len(frames))]
            digest = hashlib.sha256("".join(frames).encode()).hexdigest() if rank == 0 else None
            want = golden.get("batch")
        if rank == 0:
            parity = {"frame_sha256": digest, "oracle_sha256": want, "matches_oracle": (digest == want) if want else None,
                      "what": ("SHA-256 of the colour bytes of the frame assembled on rank 0" if cameras is None else
                               f"SHA-256 over the {len(cameras)} frame digests of one more batch" + (", rendered like the timed ones (all frames in flight, one wait)" if world > 1 and args.present == "peer" else "")) +
                              ", rendered after the timed region; oracle digest from tests/golden/bench_frames_sha256.json"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (tile kernel): algorithmic bytes per launch / CUDA-event time ----
    peak, peak_src = measured_peak_hbm()
    # bytes_tile = (Bc + Bd)*W*H + R*P + Tex, with P = (primitive, tile) pairs and R = the bytes of the
    # per-primitive record the tile kernel reads: a 4-byte bin entry plus, for indexed draws, three
    # 4-byte indices and three 16-byte post-transform vertices (64 B), else the 48-byte setup-cache record (52 B)
    pairs = last["bin_pairs"] + last["big_primitives"]
    rec_bytes = 4 + (12 + 48 if scene.index_data is not None else 48)
    tex_bytes = sum(res[1].nbytes for res in scene.bindings.values() if res[0] == "texture")
    band_px = W * (row1 - row0)
    # per launch (a C5 step holds 64 launches)
    tile_bytes = (4 + (4 if scene.has_depth else 0)) * band_px + tex_bytes + rec_bytes * pairs // passes_per_step
    tile_launch_ms = tile_ms / passes_per_step
    achieved = tile_bytes / (tile_launch_ms * 1e-3) / 1e9 if tile_ms > 0 else 0.0
    traffic, note = None, "the kernel is instruction-issue bound; HBM is the roofline the path is held to"
    try:   # one `ncu --set full` capture of this kernel on this workload (tools/make_profiles.py)
        if args.config == "c3" and world == 1:
            import glob
            latest = sorted(glob.glob(os.path.join(ROOT, "profiles", "r[0-9][0-9]_ncu_full_c3.json")))[-1]     # the latest round's capture
            for k in json.load(open(latest)):
                if k["Kernel Name"] == "wgb_tile_kernel":
                    num = lambda key: float(k[key].split()[0].replace(",", ""))
                    traffic = int((num("dram__bytes_read.sum") + num("dram__bytes_write.sum")) * 1e6)
                    note = (f"the kernel is instruction-issue bound (ncu: {num('smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f}% issue-active, "
                            f"{num('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f}% DRAM throughput; DRAM traffic below the algorithmic bytes "
                            "= vertices shared by neighbouring triangles hit L2); HBM is the roofline the path is held to")
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "kernel": "wgb_tile_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "note": note,
                "algorithmic_bytes_per_launch": int(tile_bytes), "avg_launch_ms": tile_launch_ms,
                "frame_algorithmic_bytes": int(scene.algorithmic_bytes()),
                "frame_hbm_frac": scene.algorithmic_bytes() * passes_per_step / (dev_ms * 1e-3) / 1e9 / peak if dev_ms > 0 else None}

    if args.config == "c4":
        # C4 is FP32-ALU bound.  shaders/procedural.wgsl: 12 flops per iteration x 64 iterations + 12 outside the
        # loop per pixel; no FMA contraction is allowed (bit-exactness), so the peak is one FP32 operation per lane and clock
        flops_px = 12 * 64 + 12
        sm_clock = 1.965e9
        alu_peak = 148 * 128 * sm_clock / 1e12
        alu = flops_px * W * (row1 - row0) / (tile_launch_ms * 1e-3) / 1e12 if tile_ms > 0 else 0.0
        roofline["alu"] = {"bound": "fp32 (no FMA)", "achieved": alu, "peak": alu_peak, "unit": "TFLOP/s", "frac": alu / alu_peak,
                           "flops_per_pixel": flops_px, "peak_source": "148 SMs x 128 lanes x 1.965 GHz x 1 op (FMA contraction is off)"}

    cb = None
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_baseline(scene)     # C5: one frame of the batch (the frames differ only in the camera)

    line = {
        "metric": "Mtri/s", "value": prims * args.steps / dt / 1e6, "unit": "Mtri/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "software model dry run" if model else "synthetic",
        "config": {"workload": workload, "width": W, "height": H, "triangles": prims,
                   "l2": "inputs (280 MB vertex+index) and attachments (66 MB) exceed the 126 MB L2 at C3; no explicit flush",
                   "timed_span": "K steps of submit + poll(Wait) with step k+1 submitted before step k is waited for (execution, as the reference's own pass timer); command buffers recorded before the timed region",
                   "parallelism": (f"sort-first x{world}, bands presented to " + ("rank f mod N (frame f of the batch; rank r renders the batch starting at frame r, so the ranks are on N different frames and presenters at any moment) by " if batch and args.present == "peer" else "rank 0 by ") +
                                   ("NVLink peer stores from the tile kernel" if args.present == "peer" else "NCCL send/recv"))
                   if world > 1 else "single GPU",
                   "shaders": "WGSL translated to CUDA C++ and compiled with NVRTC for sm_100a"},
        "fragments_mpix_s": last["fragments"] * args.steps / dt / 1e6,
        "shaded_mpix_s": last["shaded"] * args.steps / dt / 1e6,
        "framebuffer_mpix_s": W * H * passes_per_step * args.steps / dt / 1e6,
        "timing": {"method": "CUDA events on the render stream around the K steps (wgb_device_timer_begin / _end), max over ranks"
                             if event_ms else "wall clock between barrier + synchronize (the events failed)",
                   "event_ms_per_step": event_ms / args.steps if event_ms else None, "wall_ms_per_step": wall_dt / args.steps * 1e3},
        "device_ms_per_step": dev_ms, "geometry_ms": geom_ms, "tile_ms": tile_ms, "per_rank": per_rank,
        "pass_stats": {k: last[k] for k in ("primitives", "fragments", "shaded", "bin_pairs", "hiz_culled", "big_primitives",
                                            "clipped_primitives", "clip_records", "kernel_launches", "replays")},
        "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "parity": parity, "clocks": clocks,
        "gpu_launches": int(last["kernel_launches"]) * args.steps,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and parity["matches_oracle"] is False:
        print(f"bench.py: the rendered frame differs from the oracle's ({parity['frame_sha256']} != {parity['oracle_sha256']})", file=sys.stderr)
        sys.exit(3)


if __name__ == "__main__":
    main()
