#!/bin/bash
# usage: tools/_gpurun_retry.sh <logfile> <gpurun args...>   -- retries while the pod answers busy (exit 3 / transient)
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient" "$log" || [ $rc -eq 3 ]; then sleep 60; continue; fi
  exit $rc
done
