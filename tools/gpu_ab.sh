#!/bin/bash
# A/B of kernel tuning knobs on a B200 in one gpurun call:
#   tools/gpu.sh --timeout 1500 -- 'bash tools/gpu_ab.sh <tag> [--tests "<pytest args>"] -- "<WGB_TUNE 1>" "<WGB_TUNE 2>" ...'
# For every variant (the empty string = the default build): bench.py C3, 30 steps, no CPU baseline, no e2e ->
# gpurun_out/<tag>_ab_<n>.json (the variant is recorded in the file).  Then the launch list and one ncu full-set capture
# of the default build, and the other configs.
set -u
tag=$1; shift
tests=""
if [ "${1:-}" = "--tests" ]; then tests=$2; shift 2; fi
[ "${1:-}" = "--" ] && shift
mkdir -p gpurun_out
if [ -n "$tests" ]; then
    timeout 900 python -m pytest $tests -q -m gpu -x -p no:cacheprovider --durations=12 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/${tag}_pytest.log
fi
n=0
for v in "" "$@"; do
    if [ "${v#ENV }" != "$v" ]; then      # "ENV NAME=value": an environment variable instead of a kernel #define
        env "${v#ENV }" python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ab_${n}.json 2>> gpurun_out/${tag}_ab.err
    else
        WGB_TUNE="$v" python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ab_${n}.json 2>> gpurun_out/${tag}_ab.err
    fi
    python - "$v" gpurun_out/${tag}_ab_${n}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    d["tune"] = sys.argv[1]
    json.dump(d, open(sys.argv[2], "w"))
    p = d["pass_stats"]
    print(f"[{sys.argv[1] or 'default'}] ms/step {d['ms_per_step']:.4f} geometry {d['geometry_ms']:.4f} tile {d['tile_ms']:.4f} fragments {p['fragments']} hiz {p['hiz_culled']} parity {d.get('parity')}")
except Exception as e:
    print(f"[{sys.argv[1]}] failed: {e}")
PY
    n=$((n+1))
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
(cd wgpu-cpu_b200/csrc && ncu --set full --import-source on --clock-control none -k regex:wgb_ -s 12 -c 4 -f -o ../../gpurun_out/prof_final \
    python ../../bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1)
cp gpurun_out/${tag}_ab_0.json gpurun_out/bench_final.json
for c in c1 c2 c4 c5; do
    python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${c}.json 2>> gpurun_out/${tag}_ab.err
    python -c "
import json;d=json.loads(open('gpurun_out/bench_${c}.json').read().strip().splitlines()[-1]);print('${c}', 'ms/step %.4f geometry %.4f tile %.4f'%(d['ms_per_step'],d['geometry_ms'],d['tile_ms']), 'parity', d.get('parity'))"
done
tail -5 gpurun_out/${tag}_ab.err
