#!/bin/bash
# quick A/B of kernel #defines (WGB_TUNE) on one GPU: bench.py C3 only, no profiles:  bash tools/gpu_ab_quick.sh <tag> [--config cN] "<tune 1>" ...
set -u
tag=$1; shift
cfg=c3
if [ "${1:-}" = "--config" ]; then cfg=$2; shift 2; fi
mkdir -p gpurun_out
n=0
for v in "" "$@"; do
    WGB_TUNE="$v" python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_q_${n}.json 2>> gpurun_out/${tag}_q.err
    python - "$v" gpurun_out/${tag}_q_${n}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); d["tune"] = sys.argv[1]; json.dump(d, open(sys.argv[2], "w"))
    print(f"[{sys.argv[1] or 'default'}] ms/step {d['ms_per_step']:.4f} geometry {d['geometry_ms']:.4f} tile {d['tile_ms']:.4f} parity {(d.get('parity') or {}).get('matches_oracle')}")
except Exception as e:
    print(f"[{sys.argv[1]}] failed: {e}")
PY
    n=$((n+1))
done
tail -3 gpurun_out/${tag}_q.err
