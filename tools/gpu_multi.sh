#!/bin/bash
# Multi-GPU validation on N GPUs of one box:  tools/gpu.sh --gpus N --timeout 1200 -- 'bash tools/gpu_multi.sh N <tag>'
# hardware parity test of the peer presenter, then bench.py under torchrun (C3 with both presenters, C4, one C5 batch);
# every line carries the oracle digest check of the assembled frame.
set -u
N=$1; tag=$2
mkdir -p gpurun_out
python -m pytest tests/test_multigpu_gpu.py -q -m gpu -x -p no:cacheprovider > gpurun_out/${tag}_multigpu_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${tag}_multigpu_pytest.log
run() {   # name, bench args...
    name=$1; shift
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@" \
        > gpurun_out/${tag}_${name}.out 2> gpurun_out/${tag}_${name}.err
    echo "$name exit $?"
    grep '^{' gpurun_out/${tag}_${name}.out | tail -1 > gpurun_out/${tag}_${name}.json
    python - gpurun_out/${tag}_${name}.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    e = d.get("e2e") or {}
    print(f"  ms/step {d['ms_per_step']:.4f} value {d['value']:.1f} geometry {d['geometry_ms']:.4f} tile {d['tile_ms']:.4f} e2e {e.get('value')} ({e.get('ms_per_step')} ms) parity {(d.get('parity') or {}).get('matches_oracle')}")
except Exception as ex:
    print("  no line:", ex)
PY
}
run c3_peer --steps 30 --warmup 5
run c3_nccl --steps 30 --warmup 5 --present nccl --no-e2e
run c4_peer --config c4 --steps 20 --warmup 3 --no-e2e
run c5_peer --config c5 --steps 3 --warmup 1
tail -3 gpurun_out/${tag}_c3_peer.err
