#!/bin/bash
# compute-sanitizer over a slice of the GPU parity suite (the reference has no race detection; this is the
# new build's).  One gpurun call, ~10 minutes: the tools slow the kernels 10-100x, so the slice is small scenes only.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# Reports land in gpurun_out/sanitize_<tool>.log; copy the summaries worth keeping into profiles/.
set -u
mkdir -p gpurun_out
SLICE='colored_triangle or clipped or triangle_strip or test_lines or test_points or not_equal or load_op_load or early_depth or instancing or frag_depth or multiple_draws or overflow_replay'
for tool in memcheck racecheck synccheck initcheck; do
    timeout 600 compute-sanitizer --tool "$tool" --print-limit 20 --error-exitcode 0 \
        python -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "$SLICE" > "gpurun_out/sanitize_${tool}.log" 2>&1
    echo "== $tool: exit $?"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" "gpurun_out/sanitize_${tool}.log" | tail -4
done
