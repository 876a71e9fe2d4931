#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3: nothing charged).  Usage: tools/gpurun_retry.sh [gpurun args] -- 'cmd'
for attempt in $(seq 1 40); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    echo "[retry] attempt $attempt: no slot, sleeping 45 s" >&2
    sleep 45
done
exit 3
