#!/bin/bash
# The other BASELINE configurations on one GPU, one line each:  bash tools/gpu_cfgs.sh <tag> [configs...]
set -u
tag=$1; shift
mkdir -p gpurun_out
for c in "${@:-c1 c2 c4 c5}"; do
  for cc in $c; do
    python bench.py --config $cc --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_$cc.json 2>> gpurun_out/${tag}_cfgs.err
    python - $cc gpurun_out/${tag}_$cc.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(f"[{sys.argv[1]}] ms/step {d['ms_per_step']:.4f} device {d['device_ms_per_step']:.4f} geometry {d['geometry_ms']:.4f} tile {d['tile_ms']:.4f} parity {(d.get('parity') or {}).get('matches_oracle')}")
except Exception as e:
    print(f"[{sys.argv[1]}] failed: {e}")
PY
  done
done
