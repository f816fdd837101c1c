#!/bin/bash
# ncu full-set capture of the tile kernel (and the geometry stage) of another config:  bash tools/gpu_prof_cfg.sh <config> <tag>
set -u
cfg=$1; tag=$2
mkdir -p gpurun_out
(cd wgpu-cpu_b200/csrc && ncu --set full --import-source on --clock-control none -k regex:wgb_ -s 16 -c 4 -f -o ../../gpurun_out/prof_${cfg}_${tag} \
    python ../../bench.py --config $cfg --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1)
ls -la gpurun_out/prof_${cfg}_${tag}.ncu-rep
