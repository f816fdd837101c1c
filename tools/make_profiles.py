#!/usr/bin/env python
"""Turn the raw outputs of a measurement run (gpurun_out/) into the tracked files under profiles/:

    python tools/make_profiles.py <tag>

reads   gpurun_out/bench_final.json        bench.py line (not under a profiler)
        gpurun_out/bench_reference.json    bench.py --impl reference line            (optional)
        gpurun_out/launches_final.csv      ncu --metrics gpu__time_duration.sum ...  launch list
        gpurun_out/prof_final.ncu-rep      ncu --set full, one launch of every kernel of a pass
        gpurun_out/scale_<N>.json          bench.py lines at N = 2, 4, 8              (optional)
        build/ptxas.log                    ptxas -v of the cross-compiled kernels
writes  profiles/r01_bench_c3_<tag>.json, r01_bench_reference_c3.json, r01_launches_c3_<tag>.csv,
        r01_ncu_full_c3_<tag>.json (+ r01_ncu_full_c3.json, the copy bench.py reads `traffic` from),
        r01_scaling_<tag>.json, r01_ptxas_<tag>.log, r01_summary.md"""
import csv
import json
import os
import shutil
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
RND = os.environ.get("WGB_ROUND", "r02")      # file-name prefix of the round the run belongs to

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__shared_mem_per_block_static",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
]


def read_line(path):
    for line in open(path):
        line = line.strip()
        if line.startswith("{"):
            return json.loads(line)
    raise SystemExit(f"no JSON line in {path}")


def ncu_raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        k = {"Kernel Name": d["Kernel Name"]}
        for m in METRICS:
            if m in d:
                k[m] = f"{d[m]} {units[hdr.index(m)]}".strip()
        out.append(k)
    return out


def main():
    tag = sys.argv[1]
    os.makedirs(PROF, exist_ok=True)
    bench = read_line(os.path.join(OUT, "bench_final.json"))
    json.dump(bench, open(os.path.join(PROF, f"{RND}_bench_c3_{tag}.json"), "w"), indent=1)
    ref = None
    if os.path.exists(os.path.join(OUT, "bench_reference.json")):
        ref = read_line(os.path.join(OUT, "bench_reference.json"))
        json.dump(ref, open(os.path.join(PROF, f"{RND}_bench_reference_c3.json"), "w"), indent=1)
    elif os.path.exists(os.path.join(PROF, f"{RND}_bench_reference_c3.json")):
        ref = json.load(open(os.path.join(PROF, f"{RND}_bench_reference_c3.json")))
    shutil.copy(os.path.join(OUT, "launches_final.csv"), os.path.join(PROF, f"{RND}_launches_c3_{tag}.csv"))
    if os.path.exists(os.path.join(ROOT, "build", "ptxas.log")):
        shutil.copy(os.path.join(ROOT, "build", "ptxas.log"), os.path.join(PROF, f"{RND}_ptxas_{tag}.log"))
    full = ncu_raw(os.path.join(OUT, "prof_final.ncu-rep"))
    json.dump(full, open(os.path.join(PROF, f"{RND}_ncu_full_c3_{tag}.json"), "w"), indent=1)
    json.dump(full, open(os.path.join(PROF, f"{RND}_ncu_full_c3.json"), "w"), indent=1)
    scaling = {}
    for n in (2, 4, 8):
        p = os.path.join(OUT, f"scale_{n}.json")
        if os.path.exists(p):
            scaling[n] = read_line(p)
    if scaling:
        json.dump(scaling, open(os.path.join(PROF, f"{RND}_scaling_{tag}.json"), "w"), indent=1)

    # launch list: average per kernel over the launches of the timed + warm-up passes
    rows = [r for r in csv.reader(open(os.path.join(OUT, "launches_final.csv"))) if len(r) > 10]
    hdr = rows[0]
    per = defaultdict(list)
    for r in rows[1:]:
        d = dict(zip(hdr, r))
        if d["Metric Name"] == "gpu__time_duration.sum":
            v = float(d["Metric Value"].replace(",", ""))
            per[d["Kernel Name"]].append(v / 1e3 if d["Metric Unit"] in ("ns", "nsecond") else v)
    tot = sum(sum(v) / len(v) for v in per.values())

    L = []
    w = L.append
    ps = bench["pass_stats"]
    w(f"# Round {int(RND[1:])} profile summary (B200, C3: {bench['config']['triangles']} triangles, {bench['config']['width']}x{bench['config']['height']})\n")
    w(f"Sources: `{RND}_bench_c3_{tag}.json` (bench.py, not under a profiler), `{RND}_bench_reference_c3.json` (`--impl reference`), "
      f"`{RND}_launches_c3_{tag}.csv` (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache serialised launches: "
      f"compare shares), `{RND}_ncu_full_c3_{tag}.json` (`ncu --set full --clock-control none --import-source on`, one launch of each "
      f"kernel of a pass), `{RND}_ptxas_{tag}.log` (`ptxas -v` of the kernels as cross-compiled by `build()`). Earlier files in this "
      f"directory (v1, v5) are kept for the history of the round.\n")
    w("Commands (one `gpurun` call on a fresh B200, `tools/gpu.sh`):\n\n```\n"
      "python bench.py --steps 50 --warmup 5 > gpurun_out/bench_final.json\n"
      "python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json\n"
      "ncu --metrics gpu__time_duration.sum --clock-control none -c 75 --csv --log-file gpurun_out/launches_final.csv \\\n"
      "    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e\n"
      "(cd wgpu-cpu_b200/csrc && ncu --set full --import-source on --clock-control none -k regex:wgb_ -s 15 -c 5 \\\n"
      "    -o gpurun_out/prof_final python ../../bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e)\n"
      "python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\\n"
      "    bench.py --gpus N --steps 50 --warmup 5        # N = 2, 4, 8 (gpurun --gpus 8)\n"
      "python tools/make_profiles.py <tag>                 # this file\n```\n")
    w("## bench.py\n\n| | |\n|---|---|")
    w(f"| value | {bench['value']:.0f} Mtri/s ({bench['ms_per_step']:.3f} ms/pass; device {bench['device_ms_per_step']:.3f} ms = "
      f"geometry stage {bench['geometry_ms']:.3f} + tile {bench['tile_ms']:.3f}) |")
    w(f"| fragment-shader invocations | {bench['shaded_mpix_s']:.0f} Mpix/s (one per covered pixel; {ps['shaded']} per pass) |")
    w(f"| rasterised fragments | {bench['fragments_mpix_s']:.0f} Mfrag/s ({ps['fragments']} per pass after the hierarchical depth test; "
      f"the reference rasterises and shades 62008040) |")
    w(f"| (primitive, tile) pairs | {ps['bin_pairs']} binned, {ps.get('hiz_culled', 'n/a')} dropped by the hierarchical depth test |")
    e = bench.get("e2e")
    if e:
        w(f"| e2e (H2D {e['h2d_bytes_per_step']} B + D2H {e['d2h_bytes_per_step']} B per pass) | {e['value']:.0f} Mtri/s ({e['ms_per_step']:.2f} ms/pass) |")
    cb = bench.get("cpu_baseline")
    if cb:
        w(f"| CPU oracle port, {cb['cores']} thread | {cb['value']:.2f} {cb['unit']} ({cb['sample']}) |")
    if ref:
        w(f"| `--impl reference` | {ref['value']:.2f} {ref['unit']} ({ref['ms_per_step']:.0f} ms/pass) |")
    r = bench["roofline"]
    w(f"| roofline (tile kernel, HBM) | {r['achieved']:.0f} / {r['peak']:.0f} GB/s = {r['frac']:.3f}; algorithmic {r['algorithmic_bytes_per_launch']} B per launch, "
      f"ncu DRAM traffic {r['traffic']} B |")
    w(f"| clocks during the timed region | {bench['clocks']} |\n")
    if scaling:
        w("## sort-first scaling (C3, strong scaling, bands presented to rank 0 by NVLink peer stores)\n\n| GPUs | Mtri/s | ms/pass | geometry ms | tile ms |\n|---|---|---|---|---|")
        w(f"| 1 | {bench['value']:.0f} | {bench['ms_per_step']:.3f} | {bench['geometry_ms']:.3f} | {bench['tile_ms']:.3f} |")
        for n, b in sorted(scaling.items()):
            w(f"| {n} | {b['value']:.0f} | {b['ms_per_step']:.3f} | {b['geometry_ms']:.3f} | {b['tile_ms']:.3f} |")
        w("")
    w("## launch list (share of device time)\n\n| kernel | launches | avg us | share |\n|---|---|---|---|")
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1]) / len(kv[1])):
        a = sum(v) / len(v)
        w(f"| {k} | {len(v)} | {a:.1f} | {a / tot * 100:.1f}% |")
    w(f"\nbench.py's CUDA-event split for the same pass: tile / device = {bench['tile_ms'] / bench['device_ms_per_step'] * 100:.0f}%.\n")
    others = []
    for cfg, label in (("c1", "C1 teapot 512x512"), ("c2", "C2 bunny 1920x1080, textured"), ("c4", "C4 64-iteration fragment shader 7680x4320"),
                       ("c5", "C5 64 frames of the bunny at 3840x2160"), ("c4_8gpu", "C4 on 8 GPUs (sort-first)"), ("c5_8gpu", "C5 on 8 GPUs (sort-first, frame by frame)")):
        q = os.path.join(PROF, f"{RND}_bench_{cfg}_{tag}.json")
        if os.path.exists(q):
            others.append((label, read_line(q)))
    if others:
        w("## other configs (parity-test cases; `bench.py --config ...`, no CPU baseline / e2e legs)\n\n| config | ms/step | Mtri/s | framebuffer Mpix/s | geometry ms | tile ms |\n|---|---|---|---|---|---|")
        for label, b in others:
            w(f"| {label} | {b['ms_per_step']:.3f} | {b['value']:.1f} | {b['framebuffer_mpix_s']:.0f} | {b['geometry_ms']:.3f} | {b['tile_ms']:.3f} |")
        w("")
    w("## ncu --set full\n")
    for k in full:
        w(f"**{k['Kernel Name']}**: " + "; ".join(f"{m} = {k[m]}" for m in METRICS if m in k) + "\n")
    num = lambda k, m: float(k[m].split()[0].replace(",", ""))
    tile = [k for k in full if k["Kernel Name"] == "wgb_tile_kernel"]
    vert = [k for k in full if k["Kernel Name"] == "wgb_vertex_kernel"]
    if tile:
        t = tile[0]
        w(f"Reading: the tile kernel is instruction-issue bound ({num(t, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f}% issue-active, "
          f"{num(t, 'smsp__inst_executed.sum') / 1e6:.0f} M warp instructions, {num(t, 'smsp__thread_inst_executed_per_inst_executed.ratio'):.1f} active threads per instruction, "
          f"{num(t, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f}% DRAM throughput); its DRAM traffic "
          f"({(num(t, 'dram__bytes_read.sum') + num(t, 'dram__bytes_write.sum')):.0f} MB per launch) is below its algorithmic bytes ({r['algorithmic_bytes_per_launch'] / 1e6:.0f} MB) because "
          f"the post-transform vertices shared by neighbouring triangles and the bins are served from L2 -- no wasted re-reads. Shared memory: "
          f"{t['launch__shared_mem_per_block_static']} static per CTA, {num(t, 'launch__occupancy_limit_shared_mem'):.0f} CTAs/SM by shared memory and {num(t, 'launch__occupancy_limit_registers'):.0f} by registers, "
          f"{num(t, 'smsp__inst_executed_op_shared_atom.sum') / 1e6:.2f} M shared atomic instructions (the 64-bit atomicMin depth/order resolve).")
    if vert:
        v = vert[0]
        w(f"\nThe vertex kernel is at the HBM roofline: {(num(v, 'dram__bytes_read.sum') + num(v, 'dram__bytes_write.sum')):.0f} MB in {num(v, 'gpu__time_duration.sum'):.1f} us, "
          f"{num(v, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.0f}% DRAM throughput.")
    open(os.path.join(PROF, f"{RND}_summary.md"), "w").write("\n".join(L) + "\n")
    print("\n".join(L)[:3000])


if __name__ == "__main__":
    main()
