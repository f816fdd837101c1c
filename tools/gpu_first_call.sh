#!/bin/bash
# The full measurement run of a round in one gpurun invocation (about 14 minutes of box time):
#   tools/gpu.sh --timeout 1800 -- 'bash tools/gpu_first_call.sh r02'
# 1. the whole GPU suite, not stopping at the first failure;  2. smoke();  3. the headline bench and the reference arm;
# 4. the ncu launch list and one full-set capture of each kernel of a pass;  5. the other configs;
# 6. (SANITIZE=1) compute-sanitizer over a slice of the parity suite.
# Everything lands in gpurun_out/; `python tools/make_profiles.py <tag>` turns it into profiles/<round>_*.
set -u
tag=${1:-rXX}
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke exit $?"
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/${tag}_bench.err; cat gpurun_out/bench_final.json
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2>> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
(cd wgpu-cpu_b200/csrc && ncu --set full --import-source on --clock-control none -k regex:wgb_ -s 12 -c 4 -f -o ../../gpurun_out/prof_final \
    python ../../bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1)
for c in c1 c2 c4 c5; do
    python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${c}.json 2>> gpurun_out/${tag}_bench.err
done
if [ "${SANITIZE:-0}" = 1 ]; then bash tools/sanitize.sh > gpurun_out/${tag}_sanitize_summary.log 2>&1; cat gpurun_out/${tag}_sanitize_summary.log; fi
