#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3 / transient): tools/gpurun_retry.sh <gpurun args...>
for attempt in $(seq 1 30); do
    out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
    rc=$?
    if echo "$out" | grep -q "status=transient\|retry in a few minutes" || [ $rc -eq 3 ]; then
        echo "[retry] attempt $attempt: busy, sleeping 90 s" >&2
        sleep 90
        continue
    fi
    echo "$out"
    exit $rc
done
echo "[retry] gave up"; exit 3
