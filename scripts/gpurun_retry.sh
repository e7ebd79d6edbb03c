#!/bin/bash
# gpurun with retries while the pod answers "no slot" (exit code 3): scripts/gpurun_retry.sh <logfile> <gpurun args...>
LOG=$1; shift
for i in $(seq 1 30); do
  gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if ! grep -q "status=transient" "$LOG"; then exit $rc; fi
  sleep 60
done
