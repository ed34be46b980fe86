#!/bin/bash
# usage: scripts/gpu/gpurun_retry.sh <log> <gpurun args...>   — retries while the pod answers "transient"/busy
log=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" "$log" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
exit $rc
