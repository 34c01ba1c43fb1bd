#!/bin/bash
# gpurun with retries while the pod answers "busy / no box" (exit code 3, nothing charged).
# Usage: scripts/gpurun_retry.sh LOGFILE [gpurun args...] -- 'command'
LOG=$1; shift
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$LOG"; then echo "done rc=$rc attempt=$attempt" >> "$LOG"; exit $rc; fi
  sleep 60
done
echo "gave up" >> "$LOG"
