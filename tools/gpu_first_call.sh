#!/bin/bash
# The first GPU call of a round, in one gpurun invocation (about 16 minutes of box time):
#   tools/gpu.sh --timeout 1800 -- 'bash tools/gpu_first_call.sh r02'
# 1. the whole GPU suite, not stopping at the first failure (tests/test_zz_model_verified_gpu.py has only run on the
#    software model so far);  2. the headline bench and the reference arm;  3. the ncu launch list and one full-set
#    capture of each kernel of a pass;  4. the other configs;  5. compute-sanitizer over a slice of the parity suite.
# Everything lands in gpurun_out/; `python tools/make_profiles.py <tag>` turns it into profiles/<tag>_*.
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
# A/B of the tuning experiments that are compiled in on request (each bit-exact on the model, none measured yet)
WGB_VARY_CACHE=1 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_exp_varycache.json 2>> gpurun_out/${tag}_bench.err
WGB_VARY_CACHE=1 python bench.py --config c2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_exp_varycache_c2.json 2>> gpurun_out/${tag}_bench.err
for hp in 2 4 8; do
    WGB_HIZ_PAIRS=$hp python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_exp_hizpairs${hp}.json 2>> gpurun_out/${tag}_bench.err
done
WGB_HIZ_PAIRS=4 WGB_VARY_CACHE=1 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_exp_hizpairs4_varycache.json 2>> gpurun_out/${tag}_bench.err
for c in c1 c2 c4 c5; do
    python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${c}.json 2>> gpurun_out/${tag}_bench.err
done
if [ "${SANITIZE:-0}" = 1 ]; then bash tools/sanitize.sh > gpurun_out/${tag}_sanitize_summary.log 2>&1; cat gpurun_out/${tag}_sanitize_summary.log; fi
