#!/bin/bash
# tools/gpu.sh with retries while the pod has no free slot (gpurun exit code 3: nothing is charged):
#   tools/gpu_retry.sh <log file> [gpurun args] -- '<command>'
log=$1; shift
for i in $(seq 1 30); do
    tools/gpu.sh "$@" > "$log" 2>&1
    rc=$?
    if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
    sleep 90
done
exit 3
