#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <timeout_s> [--gpus N] <command> — retries while the pod answers "busy / transient" (exit 3)
log=$1; shift; to=$1; shift
flags=()
if [ "$1" = "--gpus" ]; then flags=(--gpus "$2"); shift 2; fi
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to "${flags[@]}" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
