#!/bin/bash
# Rebuild the in-tree native libraries (they travel to the GPU box with the snapshot), then run
# a command on a B200 through gpurun:   tools/gpu.sh [--gpus N] [--timeout S] -- '<command>'
set -e
cd "$(dirname "$0")/.."
python wgpu-cpu_b200/build.py >/dev/null
make -s -C oracle liboracle.so >/dev/null
exec /usr/local/graft/bin/gpurun "$@"
