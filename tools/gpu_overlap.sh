#!/bin/bash
# A/B of the two-stream schedule (geometry stage of pass k+1 beside the tile kernel of pass k) on one GPU:
#   bash tools/gpu_overlap.sh <tag>      -> whole frame and one band of eight, with / without the overlap and the stream priorities
set -u
tag=$1
mkdir -p gpurun_out
for mode in "" "WGB_NO_STREAM_PRIORITY=1" "WGB_NO_OVERLAP=1"; do
    for band in "" "3/8"; do
        env $mode WGB_BENCH_BAND="$band" python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ov.json 2>> gpurun_out/${tag}_ov.err
        python - "$mode" "$band" gpurun_out/${tag}_ov.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])
    print(f"[{sys.argv[1] or 'overlap + priority'}] band {sys.argv[2] or 'all'}: ms/step {d['ms_per_step']:.4f} (waited: geometry {d['geometry_ms']:.4f} tile {d['tile_ms']:.4f}) parity {(d.get('parity') or {}).get('matches_oracle')}")
except Exception as e:
    print(f"[{sys.argv[1]}] failed: {e}")
PY
    done
done
tail -3 gpurun_out/${tag}_ov.err
