#!/bin/bash
# A/B of the two-stream schedule on one GPU (C3 whole / one band of eight, C1, C2, C5):  bash tools/gpu_overlap.sh
run() { # label, env...
  label=$1; shift
  for cfg in "c3 " "c3 3/8" "c1 " "c2 " "c5 "; do
    c=${cfg%% *}; band=${cfg#* }
    env "$@" WGB_BENCH_BAND="$band" python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/diag.json 2> gpurun_out/diag.err
    python - "$label" "$c $band" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/diag.json").read().strip().splitlines()[-1])
print(f"[{sys.argv[1]}] {sys.argv[2]}: ms/step {d['ms_per_step']:.4f} parity {(d.get('parity') or {}).get('matches_oracle')}")
PY
  done
}
run "default" A=1
run "default (again)" A=1
run "priority: equal" WGB_STREAM_PRIORITY=equal
run "no overlap" WGB_NO_OVERLAP=1
