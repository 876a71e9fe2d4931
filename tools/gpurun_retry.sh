#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3: nothing charged) or while an earlier call of this repo is
# still winding down.  Usage: tools/gpurun_retry.sh [gpurun args] -- 'cmd'
for attempt in $(seq 1 60); do
    out=$(mktemp)
    /usr/local/graft/bin/gpurun "$@" 2>&1 | tee "$out"
    rc=${PIPESTATUS[0]}
    if grep -q "already running" "$out"; then rm -f "$out"; echo "[retry] attempt $attempt: earlier call still running, sleeping 30 s" >&2; sleep 30; continue; fi
    rm -f "$out"
    if [ $rc -ne 3 ]; then exit $rc; fi
    echo "[retry] attempt $attempt: no slot, sleeping 45 s" >&2
    sleep 45
done
exit 3
