#!/bin/bash
# compute-sanitizer over a slice of the GPU parity suite (the reference has no race detection; this is the
# new build's).  One gpurun call, ~8 minutes: the tools slow the kernels 10-100x, so the slice is small scenes only.
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tools/sanitize.sh'
# Reports land in gpurun_out/sanitize_<tool>.log; tools/make_profiles.py copies the summaries into profiles/.
set -u
mkdir -p gpurun_out
SLICE=${SANITIZE_SLICE:-'colored_triangle or clipped or triangle_strip or test_lines or test_points or not_equal or load_op_load or early_depth or instancing or frag_depth or multiple_draws or overflow_replay'}
LIMIT=${SANITIZE_LIMIT:-240}
for tool in memcheck racecheck synccheck initcheck; do
    timeout "$LIMIT" compute-sanitizer --tool "$tool" --print-limit 20 --error-exitcode 0 \
        python -m pytest tests/test_parity_gpu.py tests/test_async_gpu.py -q -m gpu -x -p no:cacheprovider -k "$SLICE or frames_in_flight or overflow_in_the_middle" > "gpurun_out/sanitize_${tool}.log" 2>&1
    echo "== $tool: exit $?"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" "gpurun_out/sanitize_${tool}.log" | tail -4
done
