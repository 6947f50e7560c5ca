#!/bin/bash
# gpurun with retries on "no box / slot free right now" (exit code 3, nothing charged).
#   tools/gpurun_retry.sh [gpurun flags] -- '<command>'
for attempt in $(seq 1 30); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    echo "[gpurun_retry] attempt $attempt answered busy, sleeping 90 s" >&2
    sleep 90
done
exit 3
