#!/bin/bash
# The closing measurement of a round when the GPU minutes do not allow the whole suite again (about 3 minutes of box time):
#   tools/gpu.sh --timeout 600 -- 'bash tools/gpu_final_short.sh <tag>'
# smoke(), the headline bench and the reference arm, the ncu launch list and one full-set capture of each kernel of a pass.
# `python tools/make_profiles.py <tag>` turns gpurun_out/ into profiles/<round>_*.
set -u
tag=${1:-rXX}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke exit $?"
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"; cut -c1-400 gpurun_out/bench_final.json
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2>> gpurun_out/${tag}_bench.err; echo "reference exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 75 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1; echo "launch list exit $?"
(cd wgpu-cpu_b200/csrc && ncu --set full --import-source on --clock-control none -k regex:wgb_ -s 15 -c 5 -f -o ../../gpurun_out/prof_final \
    python ../../bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1); echo "full set exit $?"
