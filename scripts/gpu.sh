#!/bin/bash
# usage: scripts/gpu.sh <timeout_s> <command...>   -- retries gpurun while the pod answers "busy" (exit 3), up to ~40 min
T=$1; shift
for i in $(seq 1 100); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
