#!/bin/bash
# compute-sanitizer over the two-GPU peer-store presenter (the tile kernels of rank 1 store their band into rank 0's
# colour target over NVLink):  tools/gpu.sh --gpus 2 --timeout 900 -- 'bash tools/sanitize_multi.sh'
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck; do
    timeout 300 compute-sanitizer --tool "$tool" --target-processes all --print-limit 30 --error-exitcode 0 \
        python -m pytest tests/test_multigpu_gpu.py -q -m gpu -x -p no:cacheprovider -k two_gpu_peer > "gpurun_out/sanitize_multi_${tool}.log" 2>&1
    echo "== $tool: exit $?"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|skipped" "gpurun_out/sanitize_multi_${tool}.log" | tail -6
done
