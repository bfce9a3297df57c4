#!/bin/bash
# usage: profiles/gpurun_retry.sh <timeout-seconds> <logfile> <command...>   (build container side)
# Retries while the pod answers "busy / draining" (exit code 3: nothing was charged).
t=$1; log=$2; shift 2
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 45
done
exit 3
