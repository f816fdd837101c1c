#!/bin/bash
# A/B of kernel #defines at N = 1 and for one rank's share of an 8-way sort-first partition on ONE GPU (WGB_BENCH_BAND):
#   bash tools/gpu_band.sh <tag> "<tune 1>" ...
set -u
tag=$1; shift
mkdir -p gpurun_out
for v in "" "$@"; do
    for band in "" "3/8"; do
        WGB_TUNE="$v" WGB_BENCH_BAND="$band" python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_band.json 2>> gpurun_out/${tag}_band.err
        python - "$v" "$band" gpurun_out/${tag}_band.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])
    print(f"[{sys.argv[1] or 'default'}] band {sys.argv[2] or 'all'}: ms/step {d['ms_per_step']:.4f} geometry {d['geometry_ms']:.4f} tile {d['tile_ms']:.4f} parity {(d.get('parity') or {}).get('matches_oracle')}")
except Exception as e:
    print(f"[{sys.argv[1]}] failed: {e}")
PY
    done
done
tail -3 gpurun_out/${tag}_band.err
