#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <timeout_s> <command...> — retries while the pod answers "busy / transient" (exit 3)
log=$1; shift; to=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
