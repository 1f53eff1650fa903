#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> <command...>   -- retries while the pod answers "busy"
T=$1; shift
for i in $(seq 1 40); do
  out=$(timeout $((T + 2400)) /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
  echo "$out"
  exit 0
done
echo "gave up"
