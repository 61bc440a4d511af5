#!/bin/bash
# gpurun with retry while the pod is busy (exit code 3 with nothing charged).  usage: tools/gpurun_retry.sh <timeout_s> '<command>' [logfile]
# A call that ended because the box was lost (also exit code 3, but charged and possibly a strike) is NOT retried: look at it first.
T=$1; CMD=$2; LOG=${3:-gpurun_out/call.log}
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$CMD" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  if grep -q "was lost\|strike" "$LOG"; then cp "$LOG" "$LOG.lost"; exit 4; fi
  sleep 60
done
exit 3
