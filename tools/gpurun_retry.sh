#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun args...]  -- retries while the pod answers "busy" (nothing is charged then)
for attempt in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | grep -v "^+" | tail -40
  if ! echo "$out" | grep -q "status=transient"; then exit 0; fi
  sleep 150
done
exit 3
