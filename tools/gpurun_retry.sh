#!/bin/bash
# usage: tools/gpurun_retry.sh LOGFILE [gpurun args...] -- 'command'    retries while the pod answers busy/transient (exit 3)
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" "$LOG" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo "gpurun_retry: finished after $i attempt(s), rc=$rc" >> "$LOG"
