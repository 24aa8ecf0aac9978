#!/bin/bash
# usage: tools/gpu_retry.sh [gpurun options] -- 'command'   -- retries while the pod answers busy / transient (nothing charged)
for attempt in $(seq 1 40); do
    out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
    rc=$?
    echo "$out" | tail -40
    if echo "$out" | grep -q "status=transient\|nothing was charged\|no box or slot"; then
        echo "[gpu_retry] attempt $attempt: busy, sleeping 60 s"
        sleep 60
        continue
    fi
    exit $rc
done
exit 3
