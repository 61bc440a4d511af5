#!/bin/bash
# gpurun with retry while the pod is busy (exit code 3 = nothing charged).  usage: tools/gpurun_retry.sh <timeout_s> '<command>' [logfile]
T=$1; CMD=$2; LOG=${3:-gpurun_out/call.log}
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$CMD" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
