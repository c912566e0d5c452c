#!/bin/bash
# usage: gpu_retry.sh <logfile> <gpurun args...>   -- retries while the pod answers busy / transient
log=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" "$log" || [ $rc -eq 3 ]; then sleep 45; continue; fi
  break
done
