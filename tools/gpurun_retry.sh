#!/bin/bash
# usage: tools/gpurun_retry.sh <log file> <gpurun args...>: repeats a gpurun call while the pod answers "busy / draining" (nothing charged)
log=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
    /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
    if ! grep -q "status=transient\|retry in a few minutes\|no box\|rc=3" "$log"; then exit 0; fi
    sleep 150
done
exit 3
