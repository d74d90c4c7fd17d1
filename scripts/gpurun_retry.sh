#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3 / status transient): nothing is charged
# for those. Usage: scripts/gpurun_retry.sh [gpurun options] -- 'command'
for attempt in $(seq 1 20); do
    out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
    if echo "$out" | grep -q "status=transient"; then
        sleep 150
        continue
    fi
    echo "$out"
    exit 0
done
echo "gpurun: still busy after 20 attempts"
exit 3
