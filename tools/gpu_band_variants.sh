#!/bin/bash
# A/B of kernel #defines for chosen bands of an 8-way sort-first partition on ONE GPU (WGB_BENCH_BAND):
#   bash tools/gpu_band_variants.sh <tag> "<band> <band> ..." "<tune 1>" ...      (band "all" = the whole frame)
set -u
tag=$1; bands=$2; shift 2
mkdir -p gpurun_out
for v in "" "$@"; do
    for band in $bands; do
        b=$band; [ "$band" = all ] && b=""
        WGB_TUNE="$v" WGB_BENCH_BAND="$b" python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_band.json 2>> gpurun_out/${tag}_band.err
        python - "$v" "$band" gpurun_out/${tag}_band.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])
    s = d["pass_stats"]
    print(f"[{sys.argv[1] or 'default'}] band {sys.argv[2]}: ms/step {d['ms_per_step']:.4f} geometry {d['geometry_ms']:.4f} tile {d['tile_ms']:.4f} clipped {s['clipped_primitives']} fragments {s['fragments']} hiz {s['hiz_culled']} parity {(d.get('parity') or {}).get('matches_oracle')}")
except Exception as e:
    print(f"[{sys.argv[1]}] failed: {e}")
PY
    done
done
tail -3 gpurun_out/${tag}_band.err
